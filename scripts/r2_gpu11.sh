#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( PTK_K3_MINB=6 timeout 200 python scripts/k3_probe.py 2>&1 | grep PTK_K3 | sed "s/^/none minb=6 /"; PTK_K3_MINB=1 timeout 200 python scripts/k3_probe.py 2>&1 | grep PTK_K3 | sed "s/^/none minb=1 /" ) > gpurun_out/k3_probe5.txt
cat gpurun_out/k3_probe5.txt
( PTK_GEMM_EXACT=1 timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6 ) > gpurun_out/split_probe8.txt 2>&1
cat gpurun_out/split_probe8.txt
( timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new8.log 2>&1
tail -3 gpurun_out/pytest_new8.log
( timeout 900 python bench.py --steps 20 --warmup 5 --skip cfg5,k1,metric --no-cpu-baseline ) > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench11.json'))
print("value", d["value"], "roofline", d["roofline"]["frac"])
for k in ("cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error") if v.get(a) is not None}, (v.get("parity") or {}).get("ok"))
P
