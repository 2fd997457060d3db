#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( for mb in 4 3; do PTK_K3_MINB=$mb timeout 200 python scripts/k3_probe.py; done; PTK_K3_TPR=256 timeout 200 python scripts/k3_probe.py ) 2>&1 | grep PTK_K3 > gpurun_out/k3_probe2.txt
cat gpurun_out/k3_probe2.txt
( PTK_GEMM_EXACT=1 timeout 300 python scripts/gemm_split_probe.py 6 2>&1 | tail -6 ) > gpurun_out/split_probe5.txt 2>&1
cat gpurun_out/split_probe5.txt
( timeout 1500 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_scan.py tests/test_gpu_blas.py tests/test_gpu_random.py tests/test_gpu_careduce.py tests/test_gpu_elemwise.py tests/test_gpu_vm.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_new5.log 2>&1
tail -30 gpurun_out/pytest_new5.log
# small-GEMM launch anatomy: per-launch durations of the recurrent-step kernel, then one full capture
T=12 PREC=bf16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_cfg4mm_bf16.csv python scripts/prof_workload.py cfg4mm 1 > /dev/null 2>&1
T=6 PREC=bf16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_pair -c 2 -o gpurun_out/r2_prof_gemm_step python scripts/prof_workload.py cfg4mm 1 > gpurun_out/ncu_gemm_step.log 2>&1
ncu -i gpurun_out/r2_prof_gemm_step.ncu-rep --page raw --csv > gpurun_out/r2_prof_gemm_step_raw.csv 2>/dev/null
( timeout 1200 python bench.py --steps 20 --warmup 5 --skip cfg5,k1 ) > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench exit $?"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench6.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["per_launch_event_pair"]["frac"])
for k in ("metric_graph","cfg3","cfg4"):
    for kk,v in d.get(k,{}).items():
        if isinstance(v,dict): print(k,kk,{a:v.get(a) for a in ("ms","tflops","evals_per_s","error","cuda_graph_replay","hits","misses") if v.get(a) is not None}, json.dumps(v.get("parity"))[:400])
P
tail -5 gpurun_out/bench6.err
ls -la gpurun_out | tail -8
