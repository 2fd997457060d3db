#!/usr/bin/env bash
# N-GPU pass: the driver's launch line for bench.py, plus the collective tests on the same ranks.
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt
( timeout 600 python -m pytest tests/test_gpu_collective.py -q -m gpu --timeout 300 ) > gpurun_out/pytest_collective_n$N.log 2>&1; tail -3 gpurun_out/pytest_collective_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
tail -5 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = int(sys.argv[1])
d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "scaling", d["scaling"], "parity", json.dumps(d.get("parity"))[:400])
print("base1", d.get("weak_scaling_base_1gpu"), "selftest", d.get("collective_selftest"))
print("strong", d.get("strong_scaling"), "e2e", d.get("e2e"))
print("alts", json.dumps(d.get("collective_alternatives"))[:500])
print("roofline", d.get("roofline"))
print("others", json.dumps(d.get("others"))[:400])
PY
