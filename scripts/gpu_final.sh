#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "roofline", d["roofline"]["frac"])
for k, v in d.get("others", {}).items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
print("sharded", d.get("sharded_logp"))
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:ptk_scan -c 2 -o gpurun_out/prof_scan_full_trace -f python scripts/prof_workload.py cfg4full 1 > gpurun_out/ncu_scan_full.log 2>&1; tail -2 gpurun_out/ncu_scan_full.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo done
