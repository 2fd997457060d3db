#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 1500 python bench.py ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
tail -4 gpurun_out/bench_final.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
