#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( PTK_GEMM_EXACT=1 timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6; timeout 200 python scripts/gemm_bench.py 2>&1 | tail -4 ) > gpurun_out/split_probe7.txt 2>&1
cat gpurun_out/split_probe7.txt
( timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py tests/test_gpu_blas.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new7.log 2>&1
tail -4 gpurun_out/pytest_new7.log
( timeout 1200 python bench.py --steps 20 --warmup 5 --skip cfg5,k1 ) > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench8.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["per_launch_event_pair"]["frac"])
for k in ("metric_graph","cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error","cuda_graph_replay","hits","misses") if v.get(a) is not None}, json.dumps(v.get("parity"))[:500])
P
tail -5 gpurun_out/bench8.err
