"""GPU suite: replay the committed, pytensor-free lowered programs (tests/golden/*.program.pkl) on the device and
compare with the committed outputs of the reference C linker (tests/golden/*.npz).  This path needs neither the
reference nor the host framework on the GPU box: torch + libptk only."""

import glob
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.basename(p)[: -len(".program.pkl")] for p in glob.glob(os.path.join(GOLD, "*.program.pkl")))


@pytest.mark.parametrize("name", NAMES)
def test_program_replay_matches_reference_golden(gpu, name):
    from pytensor_b200.vm.vm import Executor, outputs_to_host

    with open(os.path.join(GOLD, name + ".program.pkl"), "rb") as fh:
        prog = pickle.load(fh)
    z = np.load(os.path.join(GOLD, name + ".npz"))
    ins = [z[f"in{k}"] for k in range(len([k for k in z.files if k.startswith("in")]))]
    exp = [z[f"out{k}"] for k in range(len([k for k in z.files if k.startswith("out")]))]
    ex = Executor(prog, allow_gc=True, use_graph=True)
    for _ in range(3):  # eager, capture, replay
        got = outputs_to_host(ex.run([np.array(a, copy=True) for a in ins]))
        for g, e in zip(got, exp):
            g = np.asarray(g)
            assert g.shape == e.shape and g.dtype == e.dtype
            if e.dtype.kind in "biu":
                np.testing.assert_array_equal(g, e)
            elif e.dtype == np.float32:
                np.testing.assert_allclose(g, e, rtol=2e-5, atol=2e-5)
            else:
                np.testing.assert_allclose(g, e, rtol=1e-8, atol=1e-9)
