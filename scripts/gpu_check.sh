#!/usr/bin/env bash
# One gpurun call: parity tests, smoke, bench, ncu launch list. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_gemm_tc.py ${PYTEST_ARGS:-} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 400 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu --timeout 90 ) > gpurun_out/pytest_gemm_tc.log 2>&1
echo "pytest gemm_tc exit $?" >> gpurun_out/pytest_gemm_tc.log
tail -5 gpurun_out/pytest_gemm_tc.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
( timeout 900 python bench.py --steps 30 --warmup 6 ${BENCH_ARGS:-} ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 4000 gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 3 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ptk_ew -s 6 -c 2 -o gpurun_out/prof_ew -f \
     python bench.py --steps 3 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
echo done
