#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONPATH=$PWD
( time timeout 900 python -m pytest tests/test_gpu_blas.py tests/test_gpu_glue.py tests/test_gpu_vm.py -q -m gpu --timeout 300 ) > gpurun_out/pytest_gpu_part.log 2>&1
echo "pytest(part) exit $?" | tee -a gpurun_out/pytest_gpu_part.log
timeout 300 python scripts/prof_workload.py cfg5 10 > gpurun_out/cfg5_eager.log 2>&1; tail -1 gpurun_out/cfg5_eager.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg5.csv \
     python scripts/prof_workload.py cfg5 2 > gpurun_out/ncu_cfg5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_small\|put_rows_kernel\|take_lastaxis -c 4 -o gpurun_out/prof_cfg5_kernels -f \
   python scripts/prof_workload.py cfg5 1 > gpurun_out/ncu_full_cfg5.log 2>&1
python - <<'PY'
import numpy as np, sys, time
sys.path.insert(0, ".")
from oracle import cvm
pytensor = cvm.configure("float64")
import pytensor.tensor as pt, pytensor_b200
A = pt.dmatrix("A"); b = pt.dmatrix("b")
L = pt.linalg.cholesky(A)
f = pytensor.function([A, b], [L, pt.linalg.solve_triangular(L, b, lower=True)], mode="CUDA")
rng = np.random.default_rng(0); n = 2048
Av = rng.standard_normal((n, n)); Av = Av @ Av.T / n + np.eye(n)
bv = rng.standard_normal((n, 64))
for _ in range(5):
    t = time.perf_counter(); r = f(Av, bv); print("chol+trsm n=2048 e2e s", time.perf_counter() - t)
import scipy.linalg
t = time.perf_counter(); Lr = scipy.linalg.cholesky(Av, lower=True); print("scipy potrf s", time.perf_counter() - t)
print("max err", np.abs(r[0] - Lr).max())
PY
echo done
