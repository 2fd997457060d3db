#!/usr/bin/env bash
# Round-2 final single-GPU pass: full GPU suite, smoke, the default bench command, ncu evidence.  Everything -> gpurun_out/.
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 1500 python -m pytest tests -q -m gpu --timeout 300 --maxfail=40 ) > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log; tail -6 gpurun_out/pytest_gpu_final.log
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time timeout 1500 python bench.py ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench_final.json; tail -4 gpurun_out/bench_final.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
# ncu: launch list of the (trimmed) default command, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 3 --reps 3 --no-cpu-baseline --skip metric,cfg3,cfg4,cfg5 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ptk_ew_red_row -c 2 -o gpurun_out/r2_prof_k3_final python scripts/k3_probe.py > gpurun_out/ncu_k3_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_pair -c 2 -o gpurun_out/r2_prof_gemm_bf16 python scripts/gemm_bench.py > gpurun_out/ncu_gemm_bf16.log 2>&1
PTK_GEMM_EXACT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc_pair -c 1 -o gpurun_out/r2_prof_gemm_exact python scripts/gemm_split_probe.py 6 > gpurun_out/ncu_gemm_exact.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ptk_rowfuse -c 2 -o gpurun_out/r2_prof_cfg5_rowfuse python scripts/prof_workload.py cfg5 1 > gpurun_out/ncu_cfg5.log 2>&1
for f in r2_prof_k3_final r2_prof_gemm_bf16 r2_prof_gemm_exact r2_prof_cfg5_rowfuse; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -20
