#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 600 python -m pytest tests/test_gpu_collective.py -q -m gpu --timeout 500 ) > gpurun_out/pytest_collective.log 2>&1
echo "collective test exit $?"; tail -15 gpurun_out/pytest_collective.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 30 --warmup 6 ) > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench N=2 exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"], 1), "sharded", d.get("sharded_logp"))
PY
echo done
