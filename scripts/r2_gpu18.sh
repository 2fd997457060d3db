#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( timeout 400 python -m pytest tests/test_gpu_blas.py tests/test_gpu_vm.py -q -m gpu --timeout 200 --maxfail=30 ) > gpurun_out/pytest_chain2.log 2>&1
tail -5 gpurun_out/pytest_chain2.log
