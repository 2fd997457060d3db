#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for t in auto 64 128 256; do if [ $t = auto ]; then timeout 200 python scripts/k3_probe.py; else PTK_K3_TPR=$t timeout 200 python scripts/k3_probe.py; fi; done 2>&1 | grep PTK_K3 > gpurun_out/k3_probe.txt
cat gpurun_out/k3_probe.txt
for kc in 8; do PTK_GEMM_KCHUNK=$kc timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6; done > gpurun_out/split_probe3.txt 2>&1
cat gpurun_out/split_probe3.txt
( timeout 1200 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py tests/test_gpu_elemwise.py tests/test_gpu_careduce.py tests/test_gpu_vm.py tests/test_gpu_shared.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new3.log 2>&1
tail -30 gpurun_out/pytest_new3.log
( timeout 1200 python bench.py --steps 20 --warmup 5 --skip cfg5,k1 ) > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench4.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["whole_graph"]["frac"])
for k in ("metric_graph","cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error","cuda_graph_replay") if v.get(a) is not None}, json.dumps(v.get("parity"))[:300])
P
tail -5 gpurun_out/bench4.err
# ncu: full capture of the K3 kernel + launch list of the default bench command
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ptk_ew_red_row -c 3 -o gpurun_out/r2_prof_k3 python scripts/k3_probe.py > gpurun_out/ncu_k3.log 2>&1
ncu -i gpurun_out/r2_prof_k3.ncu-rep --page raw --csv > gpurun_out/r2_prof_k3_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
