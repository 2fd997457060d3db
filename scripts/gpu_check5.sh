#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 1800 python -m pytest tests -q -m gpu --timeout 300 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 1200 python bench.py --steps 30 --warmup 6 --extra ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
for k, v in d.get("others", {}).items(): print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
print("sharded", d.get("sharded_logp"))
PY
for w in cfg3 cfg5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$w.csv \
     python scripts/prof_workload.py $w 2 > gpurun_out/ncu_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:potrf_diag\|trsm_diag\|potrf_panel -c 3 -o gpurun_out/prof_linalg -f \
   python - > gpurun_out/ncu_full_linalg.log 2>&1 <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from oracle import cvm
pytensor = cvm.configure("float64")
import pytensor.tensor as pt, pytensor_b200
A = pt.dmatrix("A"); b = pt.dmatrix("b")
L = pt.linalg.cholesky(A)
f = pytensor.function([A, b], [L, pt.linalg.solve_triangular(L, b, lower=True)], mode="CUDA")
rng = np.random.default_rng(0); n = 2048
Av = rng.standard_normal((n, n)); Av = Av @ Av.T / n + np.eye(n)
import time
for _ in range(3):
    t = time.perf_counter(); r = f(Av, rng.standard_normal((n, 64))); print("chol+trsm n=2048 s", time.perf_counter() - t)
import scipy.linalg
t = time.perf_counter(); Lr = scipy.linalg.cholesky(Av, lower=True); print("scipy potrf s", time.perf_counter() - t)
print("max err", np.abs(r[0] - Lr).max())
PY
tail -6 gpurun_out/ncu_full_linalg.log
echo done
