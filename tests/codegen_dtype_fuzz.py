"""Random mixed-dtype elementwise graphs (signed / unsigned / bool / float32 / float64 operands): the scalar bodies this backend
GENERATES are compiled for the host (tests/test_scalar_table_cpu.py::_emulate) and compared with the reference C linker —
i.e. our C expressions against the reference's `ScalarOp.c_code` under C's usual arithmetic conversions, without a GPU.

    python tests/codegen_dtype_fuzz.py FIRST LAST
"""

import os
import pathlib
import sys
import tempfile
import warnings

import numpy as np

if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from helpers import pytensor  # noqa: E402

import pytensor.tensor as pt  # noqa: E402
import test_scalar_table_cpu as T  # noqa: E402

DTYPES = ["int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "bool", "float32", "float64"]


def build(rng):
    N = 40
    ins, vals = [], []
    for nm in "pqr":
        dt = DTYPES[int(rng.integers(len(DTYPES)))]
        ins.append(pt.vector(nm, dtype=dt))
        if dt == "bool":
            vals.append(rng.random(N) < 0.5)
        elif dt.startswith("float"):
            vals.append((rng.standard_normal(N) * 3).astype(dt))
        else:
            vals.append(rng.integers(0 if dt.startswith("u") else -6, 7, size=N).astype(dt))
    pool = list(ins)

    def pick():
        return pool[int(rng.integers(len(pool)))]

    for _ in range(int(rng.integers(1, 6))):
        k = int(rng.integers(0, 14))
        x, y = pick(), pick()
        try:
            if k == 0:
                r = x + y
            elif k == 1:
                r = x * y
            elif k == 2:
                r = x - y
            elif k == 3:
                r = pt.switch(x > y, x, y)
            elif k == 4:
                r = pt.maximum(x, y)
            elif k == 5:
                r = pt.minimum(x, y)
            elif k == 6:
                r = (x < y) | pt.eq(x, y)
            elif k == 7:
                r = pt.abs(x) if x.dtype != "bool" else ~x
            elif k == 8:
                r = pt.clip(x, -2, 3) if x.dtype != "bool" else x
            elif k == 9:
                r = x / pt.switch(pt.eq(y, 0), pt.ones_like(y), y)
            elif k == 10:
                to = DTYPES[int(rng.integers(len(DTYPES)))]
                if x.dtype.startswith("float") and to.startswith("uint"):
                    to = to[1:]      # (a negative float converted to an UNSIGNED integer is undefined behaviour in C)
                r = pt.cast(x, to)
            elif k == 11:      # float functions of integer operands (the result type is upgraded to a float)
                fn = [pt.exp, pt.tanh, pt.sigmoid, pt.sqrt, pt.log1p, pt.sin, pt.arctan, pt.erf][int(rng.integers(8))]
                r = fn(pt.abs(x) if fn in (pt.sqrt, pt.log1p) else x) if x.dtype != "bool" else fn(x)
            elif k == 12:
                r = pt.sqr(x) + pt.sgn(y) if "bool" not in (x.dtype, y.dtype) else x ^ y if x.dtype == y.dtype == "bool" else x + y
            else:
                r = pt.round(x / 3) if x.dtype.startswith("float") else pt.floor(x / 3)
            pool.append(r)
        except Exception:  # noqa: BLE001
            pass
    outs = [pick() for _ in range(int(rng.integers(1, 3)))]
    outs.append(pt.switch(ins[0] > ins[1], 1, 0) + pt.switch(ins[2] > ins[1], 1, 0))    # every input is used
    return ins, outs, vals


def check_seed(seed):
    """"ok" | "skipped" (the graph does not lower to plain elementwise steps); raises on a mismatch."""
    rng = np.random.default_rng(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            ins, outs, vals = build(rng)
            pytensor.function(ins, outs, mode="CVM")
        except Exception:  # noqa: BLE001
            return "skipped"
        try:
            got, ref = T._emulate(ins, outs, vals, pathlib.Path(tempfile.mkdtemp()))
        except AssertionError:
            return "skipped"
    f = pytensor.function(ins, outs, mode="CUDA")
    single = any(d == "float32" for st in f.vm.executor.program.steps if hasattr(st.impl, "prog")
                 for i in st.impl.prog.insts for d in list(i.in_dtypes) + [i.out_dtype])
    for g, r in zip(got, ref):
        r = np.asarray(r)
        if r.dtype == np.bool_:      # (a C `1 + 1` stored into a bool byte is 2: compare truth values)
            assert np.array_equal(g.view(np.uint8) != 0, r.view(np.uint8) != 0), seed
        elif r.dtype.kind in "iu":
            # beyond 2^53 the reference's Maximum / Minimum lose bits (the selected operand passes through a double); this
            # backend keeps 64-bit integers exact there — a deliberate deviation, not compared
            small = np.abs(r.astype(np.float64)) < 2.0 ** 53
            np.testing.assert_array_equal(g[small], r[small], err_msg=f"seed {seed}")
        else:
            # a float32 intermediate (e.g. exp of an int8: `expf((float)x)` here, `exp(x)` in double rounded to float32 there)
            # may differ in its last bit: the 1e-5 contract, not bit equality
            np.testing.assert_allclose(g, r, rtol=1e-6 if (r.dtype == np.float32 or single) else 1e-13, atol=0, equal_nan=True,
                                       err_msg=f"seed {seed}")
    return "ok"


if __name__ == "__main__":
    count = {"ok": 0, "skipped": 0, "FAILED": 0}
    for s in range(int(sys.argv[1]), int(sys.argv[2])):
        try:
            count[check_seed(s)] += 1
        except Exception as e:  # noqa: BLE001
            count["FAILED"] += 1
            print("SEED", s, type(e).__name__, str(e)[:400].replace("\n", " | "))
    print(count)
