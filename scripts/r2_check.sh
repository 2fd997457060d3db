#!/usr/bin/env bash
# One gpurun call: GPU test suite, smoke, bench (N=1). Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
if [ "${TESTS:-1}" = "1" ]; then
( time timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -q -m gpu --timeout 300 --maxfail=40 ${PYTEST_ARGS:-} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
fi
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
( time timeout 1200 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:-} ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 6000 gpurun_out/bench.json
tail -8 gpurun_out/bench.err
echo done
