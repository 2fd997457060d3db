#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 900 python -m pytest tests/test_gpu_vm.py tests/test_gpu_gemm_tc.py tests/test_gpu_elemwise.py -q -m gpu --timeout 300 ) > gpurun_out/pytest_part.log 2>&1
echo "pytest(part) exit $?" | tee -a gpurun_out/pytest_part.log
tail -5 gpurun_out/pytest_part.log
( timeout 900 python bench.py --steps 30 --warmup 6 ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "roofline", d["roofline"]["frac"])
PY
for ch in 1 4 16; do
python - <<PY
import numpy as np, sys, time
sys.path.insert(0, ".")
from oracle import cvm
pytensor = cvm.configure("float32")
import pytensor_b200
from pytensor_b200 import workloads as W
from pytensor_b200.vm.vm import Executor
import bench
Executor.STREAM_CHUNKS = $ch
ins, outs, mk, _ = W.cfg2_fused_elemwise(4096)
f = pytensor.function(ins, outs, mode="CUDA", trust_input=True)
args = [bench.pinned_like(a) for a in mk(1)]
for _ in range(3): f(*args)
t = time.perf_counter()
for _ in range(20): r = f(*args)
dt = (time.perf_counter() - t) / 20
print("chunks", $ch, "e2e ms", dt * 1e3, "chunked", f.vm.executor.chunked_calls)
PY
done
echo done
