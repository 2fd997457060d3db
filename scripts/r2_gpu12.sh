#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 200 python scripts/gemm_bench.py 2>&1 | tail -4; timeout 200 python scripts/gemm_latency_probe.py 2>&1 | tail -10 ) > gpurun_out/gemm_bench12.txt 2>&1
cat gpurun_out/gemm_bench12.txt
( timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py tests/test_gpu_blas.py tests/test_gpu_careduce.py tests/test_gpu_vm.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new9.log 2>&1
tail -3 gpurun_out/pytest_new9.log
( timeout 900 python bench.py --steps 20 --warmup 5 --skip cfg5,k1,metric --no-cpu-baseline ) > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench12.json'))
print("value", d["value"], "roofline", d["roofline"]["frac"])
for k in ("cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error") if v.get(a) is not None}, (v.get("parity") or {}).get("ok"))
P
