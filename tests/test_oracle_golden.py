"""CPU suite: pins the NumPy port oracle (oracle/numpy_port.py) against the golden vectors produced by the unmodified
reference C linker (tests/golden/*.npz), exercises the lowering of every golden graph, and writes/reads the
pytensor-free program pickles the GPU suite replays."""

import os
import pickle

import numpy as np
import pytest

from helpers import pytensor  # noqa: F401  (configures the host framework + registers mode="CUDA")

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

import sys  # noqa: E402

sys.path.insert(0, GOLD)
from cases import CASES, CPU_ONLY  # noqa: E402

from oracle import numpy_port  # noqa: E402


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    ins = [z[f"in{k}"] for k in range(len([k for k in z.files if k.startswith("in")]))]
    outs = [z[f"out{k}"] for k in range(len([k for k in z.files if k.startswith("out")]))]
    return ins, outs


def _tol(name, dtype):
    if "scan" in name or "mlp" in name or "elemwise" in name:
        return dict(rtol=2e-5, atol=2e-5)
    if "special" in name:   # the reference's own C support code (AS 103 / AS 121 with truncated coefficients, Cephes incbet) vs SciPy
        return dict(rtol=1e-7, atol=1e-9)
    return dict(rtol=1e-6, atol=1e-6) if np.dtype(dtype) == np.float32 else dict(rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("name", sorted(CASES))
def test_numpy_port_matches_reference_golden(name):
    ins_v, outs_v, args, floatX = CASES[name]()
    gin, gout = _load(name)
    for a, b in zip(args, gin):
        np.testing.assert_array_equal(np.asarray(a), b)  # the committed inputs are the seeded ones
    f = pytensor.function(ins_v, outs_v, mode="CUDA")   # lowering only (no device needed)
    prog = f.vm.executor.program
    res = numpy_port.evaluate_program(prog, gin)
    assert len(res) == len(gout)
    for r, e in zip(res, gout):
        assert r.shape == e.shape and r.dtype == e.dtype, (r.shape, e.shape, r.dtype, e.dtype)
        if e.dtype.kind in "biu":
            np.testing.assert_array_equal(r, e)
        else:
            np.testing.assert_allclose(r, e, **_tol(name, e.dtype))
    # ship the lowered program (pytensor-free) next to the fixture for the GPU suite
    blob = pickle.dumps(prog)
    prog2 = pickle.loads(blob)
    res2 = numpy_port.evaluate_program(prog2, gin)
    for r, e in zip(res2, res):
        np.testing.assert_array_equal(r, e)
    if name not in CPU_ONLY:   # (CPU_ONLY cases pin the port oracle only; the GPU suite replays what has a pickle)
        path = os.path.join(GOLD, name + ".program.pkl")
        if not (os.path.exists(path) and _same_program(path, prog, gin, res)):
            with open(path, "wb") as fh:   # (byte-different pickles of the SAME program — string sharing depends on what
                fh.write(blob)             #  ran before in the process — are not rewritten: the tree stays clean)


def _same_program(path, prog, gin, res):
    try:
        with open(path, "rb") as fh:
            old = pickle.load(fh)
        shape = lambda p: [(type(s.impl).__name__, tuple(s.ins), tuple(s.outs)) for s in p.steps]  # noqa: E731
        if shape(old) != shape(prog) or old.n_slots != prog.n_slots or list(old.inputs) != list(prog.inputs) \
                or list(old.outputs) != list(prog.outputs):
            return False
        return all(np.array_equal(a, b, equal_nan=True) for a, b in zip(numpy_port.evaluate_program(old, gin), res))
    except Exception:  # noqa: BLE001  (unreadable / outdated pickle: write a fresh one)
        return False
