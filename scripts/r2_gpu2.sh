#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for kc in 0 4 8 16; do PTK_GEMM_KCHUNK=$kc timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6; done > gpurun_out/split_probe.txt 2>&1
PTK_GEMM_KCHUNK=8 timeout 300 python scripts/gemm_split_probe.py 3 2>&1 | tail -6 >> gpurun_out/split_probe.txt
cat gpurun_out/split_probe.txt
( timeout 1200 python -m pytest tests/test_gpu_special.py tests/test_gpu_advindex.py tests/test_gpu_gemm_tc.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new.log 2>&1
tail -40 gpurun_out/pytest_new.log
( timeout 900 python bench.py --steps 20 --warmup 5 --skip cfg4,cfg5,k1 ) > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench2.json'))
print("value", d["value"], "e2e", d["e2e"])
for k in ("metric_graph","cfg3"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s")}, v.get("parity"))
P
tail -5 gpurun_out/bench2.err
