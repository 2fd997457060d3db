#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for kc in 0 8 16; do PTK_GEMM_KCHUNK=$kc timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6; done > gpurun_out/split_probe2.txt 2>&1
PTK_GEMM_KCHUNK=8 timeout 300 python scripts/gemm_split_probe.py 3 2>&1 | tail -6 >> gpurun_out/split_probe2.txt
timeout 300 python scripts/gemm_bench.py 2>&1 | tail -4 >> gpurun_out/split_probe2.txt
cat gpurun_out/split_probe2.txt
( timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py tests/test_gpu_blas.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new2.log 2>&1
tail -30 gpurun_out/pytest_new2.log
( timeout 1200 python bench.py --steps 20 --warmup 5 --skip cfg5,k1 ) > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench3.json'))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k in ("metric_graph","cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error") if v.get(a) is not None}, json.dumps(v.get("parity"))[:600])
P
tail -5 gpurun_out/bench3.err
