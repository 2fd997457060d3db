#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( for m in none tma; do PTK_K3_PIPE=$m timeout 200 python scripts/k3_probe.py 2>&1 | grep PTK_K3 | sed "s/^/pipe=$m /"; done
  PTK_K3_PIPE=tma PTK_K3_MINB=4 timeout 200 python scripts/k3_probe.py 2>&1 | grep PTK_K3 | sed "s/^/pipe=tma minb=4 /"
  PTK_K3_PIPE=tma PTK_K3_TPR=256 timeout 200 python scripts/k3_probe.py 2>&1 | grep PTK_K3 | sed "s/^/pipe=tma tpr=256 /" ) > gpurun_out/k3_probe4.txt
cat gpurun_out/k3_probe4.txt
( PTK_K3_PIPE=tma timeout 900 python -m pytest tests/test_gpu_careduce.py tests/test_gpu_elemwise.py tests/test_gpu_vm.py tests/test_gpu_golden.py -q -m gpu --timeout 300 --maxfail=30 ) > gpurun_out/pytest_tma.log 2>&1
tail -4 gpurun_out/pytest_tma.log
PTK_K3_PIPE=tma timeout 600 ncu --set full --clock-control none --import-source on -k regex:ptk_ew_red_row -c 2 -o gpurun_out/r2_prof_k3_tma python scripts/k3_probe.py > gpurun_out/ncu_k3_tma.log 2>&1
ncu -i gpurun_out/r2_prof_k3_tma.ncu-rep --page raw --csv > gpurun_out/r2_prof_k3_tma_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
