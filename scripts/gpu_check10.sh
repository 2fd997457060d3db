#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_blas.py -q -m gpu --timeout 300 ) > gpurun_out/pytest_gemm.log 2>&1
echo "pytest(gemm) exit $?" | tee -a gpurun_out/pytest_gemm.log
tail -3 gpurun_out/pytest_gemm.log
PTK_GEMM_SPLIT=0 timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_nosplit.log
PTK_GEMM_SPLIT=1 timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_split.log
for sp in 0 1; do
PTK_GEMM_SPLIT=$sp timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_bf16 -c 12 --csv --log-file gpurun_out/gemm_ncu_split$sp.csv python scripts/gemm_bench.py > /dev/null 2>&1
done
echo done
