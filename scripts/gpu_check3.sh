#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 5000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for w in cfg3 cfg5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_$w.csv \
     python scripts/prof_workload.py $w 2 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 2 -c 1 -o gpurun_out/prof_gemm_tc -f \
   python scripts/prof_workload.py cfg3 1 > gpurun_out/ncu_full_gemm.log 2>&1
N=64 timeout 300 python scripts/prof_workload.py metric 20 > gpurun_out/metric64.log 2>&1; tail -1 gpurun_out/metric64.log
echo done
