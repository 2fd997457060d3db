#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 --durations=12 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for sp in 0 1; do
PTK_GEMM_SPLIT=$sp timeout 300 python scripts/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench_split$sp.log
done
( timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"], "roofline", d["roofline"]["frac"])
for k, v in d.get("others", {}).items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
print("sharded", d.get("sharded_logp"))
PY
timeout 500 python scripts/pcie_probe.py 2>&1 | grep -v "^$" | cut -c1-160 > gpurun_out/pcie_probe.log; cat gpurun_out/pcie_probe.log
grep -A14 'slowest' gpurun_out/pytest_gpu.log
echo done
