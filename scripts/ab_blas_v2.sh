#!/usr/bin/env bash
# A/B of the restructured skinny-GEMM kernels (PTK_BLAS_V2=1; host-emulated in tests/test_kernels_cpu_emulation.py,
# not yet timed on a device): parity tests under the flag, then cfg5 (logp+grad, 2^17 chains) with and without it.
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
PTK_BLAS_V2=1 timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_vm.py tests/test_gpu_golden.py -q -m gpu --timeout 300 2>&1 | tail -3
for v in 0 1; do
  PTK_BLAS_V2=$v timeout 300 python scripts/prof_workload.py cfg5 20 2>&1 | tail -1 | sed "s/^/PTK_BLAS_V2=$v /"
done
for v in 0 1; do
  PTK_BLAS_V2=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
    -k regex:gemm_small -c 8 --csv --log-file gpurun_out/blas_v2_$v.csv python scripts/prof_workload.py cfg5 1 > /dev/null 2>&1
done
echo done
