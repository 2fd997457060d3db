#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for mode in 3 2 1; do
  ( PTK_GEMM_MODE=$mode timeout 300 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu --timeout 100 ) > gpurun_out/pytest_gemm_mode$mode.log 2>&1
  echo "gemm mode $mode exit $?" | tee -a gpurun_out/pytest_gemm_mode$mode.log
  PTK_GEMM_MODE=$mode timeout 200 python scripts/prof_workload.py cfg3 10 > gpurun_out/cfg3_mode$mode.log 2>&1; tail -1 gpurun_out/cfg3_mode$mode.log
done
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 --deselect tests/test_gpu_gemm_tc.py ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( PTK_GEMM_MODE=${BEST_MODE:-3} timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["clocks"])
for k, v in d.get("others", {}).items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
print("sharded", d.get("sharded_logp"))
PY
PTK_GEMM_MODE=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 2 -c 1 -o gpurun_out/prof_gemm_pair -f \
   python scripts/prof_workload.py cfg3 1 > gpurun_out/ncu_full_gemm_pair.log 2>&1
echo done
