#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python scripts/r2_probe1.py > gpurun_out/r2_probe1.json 2> gpurun_out/r2_probe1.err; echo "probe1 exit $?"
cat gpurun_out/r2_probe1.json; tail -3 gpurun_out/r2_probe1.err
PTK_BLAS_V2=1 timeout 600 python -m pytest tests/test_gpu_blas.py tests/test_gpu_vm.py tests/test_gpu_golden.py -q -m gpu --timeout 300 2>&1 | tail -3
for v in 0 1; do
  PTK_BLAS_V2=$v timeout 300 python scripts/prof_workload.py cfg5 20 2>&1 | tail -1 | sed "s/^/PTK_BLAS_V2=$v /"
done
echo done
