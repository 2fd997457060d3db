#!/usr/bin/env bash
# multi-GPU bench: N = $1 ranks on one box
set -u
N=${1:-2}
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 30 --warmup 6 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
( timeout 600 python bench.py --gpus 1 --steps 30 --warmup 6 --no-cpu-baseline ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - $N <<'PY'
import json, sys
for n in (1, int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    try:
        d = json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"], 1), "sharded", d.get("sharded_logp", {}).get("evals_per_s"), d.get("sharded_logp", {}).get("ms_per_eval"))
    except Exception as e:
        print(n, "ERR", e)
PY
echo done
