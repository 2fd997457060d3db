#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 400 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu --timeout 120 ) > gpurun_out/pytest_gemm_tc.log 2>&1
echo "gemm_tc(cluster2) exit $?" | tee -a gpurun_out/pytest_gemm_tc.log
( PTK_GEMM_CLUSTER=1 timeout 400 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu --timeout 120 ) > gpurun_out/pytest_gemm_tc_c1.log 2>&1
echo "gemm_tc(cluster1) exit $?" | tee -a gpurun_out/pytest_gemm_tc_c1.log
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_gemm_tc.py ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
PTK_GEMM_CLUSTER=1 timeout 300 python scripts/prof_workload.py cfg3 10 > gpurun_out/cfg3_cluster1.log 2>&1; tail -1 gpurun_out/cfg3_cluster1.log
timeout 300 python scripts/prof_workload.py cfg3 10 > gpurun_out/cfg3_cluster2.log 2>&1; tail -1 gpurun_out/cfg3_cluster2.log
for w in cfg3 cfg5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$w.csv \
     python scripts/prof_workload.py $w 2 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 2 -c 1 -o gpurun_out/prof_gemm_tc -f \
   python scripts/prof_workload.py cfg3 1 > gpurun_out/ncu_full_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptk_scan_fused -c 1 -o gpurun_out/prof_scan -f \
   python scripts/prof_workload.py cfg4 1 > gpurun_out/ncu_full_scan.log 2>&1
echo done
