#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_vm.py tests/test_gpu_gemm_tc.py tests/test_gpu_golden.py tests/test_gpu_shared.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_chain.log 2>&1
tail -5 gpurun_out/pytest_chain.log
( timeout 900 python bench.py --steps 20 --warmup 5 --skip cfg3,cfg4,cfg5,k1 --no-cpu-baseline ) > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench17.json'))
print("value", d["value"])
for kk,v in d.get("metric_graph",{}).items():
    if isinstance(v,dict): print(kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error","launches_per_eval_after_fusion") if v.get(a) is not None}, (v.get("parity") or {}).get("ok"), (v.get("parity") or {}).get("max_rel_to_scale"), v.get("cpu_reference"))
P
tail -3 gpurun_out/bench17.err
