"""General advanced indexing on the device (vm/nodes_advidx.py) against the reference C linker: boolean masks, index
arrays separated by slices (NumPy moves those axes to the front), partial slices mixed with index arrays, 0-d and
multi-dimensional index arrays, and set / inc through all of them (reference: AdvancedSubtensor
pytensor/tensor/subtensor.py:1932-2236, AdvancedIncSubtensor :2275; its tests tests/tensor/test_subtensor.py
`TestAdvancedSubtensor`).  Gathers and sets are bit-exact; increments with duplicates compare within fp tolerance."""

import os

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu


def _data(seed=41, shape=(7, 9, 11)):
    rng = np.random.default_rng(seed)
    return rng, rng.standard_normal(shape).astype("float32")


def test_boolean_masks_get(gpu):
    rng, xv = _data()
    x = pt.ftensor3("x")
    m3, m2, m1 = pt.tensor3("m3", dtype="bool"), pt.matrix("m2", dtype="bool"), pt.vector("m1", dtype="bool")
    mv3, mv2, mv1 = xv > 0.3, rng.random((7, 9)) < 0.4, rng.random(9) < 0.5
    compare_cuda_and_cvm([x, m3], [x[m3]], [xv, mv3], exact=True)
    compare_cuda_and_cvm([x, m2], [x[m2]], [xv, mv2], exact=True)
    compare_cuda_and_cvm([x, m1], [x[:, m1], x[:, m1, 2:9:3]], [xv, mv1], exact=True)
    compare_cuda_and_cvm([x], [x[x > 5.0]], [xv], exact=True)          # nothing selected
    compare_cuda_and_cvm([x], [x[pt.abs(x) >= 0.0]], [xv], exact=True)  # everything selected
    big = rng.standard_normal((300, 257)).astype("float32")              # several compaction tiles, unaligned rows
    y = pt.fmatrix("y")
    compare_cuda_and_cvm([y], [y[y > 0.5], (y * 2)[y < -1]], [big], exact=True)


def test_non_consecutive_and_mixed_index_arrays_get(gpu):
    rng, xv = _data(42)
    x = pt.ftensor3("x")
    i, j = pt.lvector("i"), pt.lvector("j")
    iv, jv = rng.integers(-7, 7, 5), rng.integers(-11, 11, 5)
    compare_cuda_and_cvm([x, i, j], [x[i, :, j], x[i, 1:8:2, j], x[i, ::-1, j]], [xv, iv, jv], exact=True)
    compare_cuda_and_cvm([x, i], [x[1:, i], x[::2, :, i], x[2, i], x[i, ::2], x[-3:, i, 1:-1]], [xv, iv], exact=True)
    I = pt.lmatrix("I")
    Iv = rng.integers(0, 7, (3, 4))
    compare_cuda_and_cvm([x, I, j], [x[I], x[:, I], x[I, :, I + 1], x[I[:, :1], 2:5, j[:4]]], [xv, Iv, jv], exact=True)
    m1 = pt.vector("m1", dtype="bool")
    mv = np.array([1, 0, 1, 1, 0, 0, 1], dtype=bool)
    j4v = np.array([0, -9, 8, 3])   # axis 1 has 9 entries; the mask selects 4 rows, the two broadcast together
    if os.environ.get("PTK_DRY") != "1":  # (the CPU dry run has no mask contents, hence no count to broadcast against)
        compare_cuda_and_cvm([x, m1, j], [x[m1, j], x[m1, :, j]], [xv, mv, j4v], exact=True)


def test_set_and_inc_through_general_indices(gpu):
    rng, xv = _data(43)
    x = pt.ftensor3("x")
    m3 = pt.tensor3("m3", dtype="bool")
    i, j = pt.lvector("i"), pt.lvector("j")
    iv, jv = np.array([0, 3, 6, 2, 5]), np.array([1, 4, 7, 10, 0])   # distinct (i, j) pairs: sets are well defined
    mv = xv < 0
    compare_cuda_and_cvm([x, m3], [pt.set_subtensor(x[m3], 0.0), pt.inc_subtensor(x[m3], 1.5)], [xv, mv], exact=True)
    v = pt.fmatrix("v")
    vv = rng.standard_normal((5, 9)).astype("float32")
    compare_cuda_and_cvm([x, i, j, v], [pt.set_subtensor(x[i, :, j], v), pt.inc_subtensor(x[i, :, j], v)], [xv, iv, jv, vv],
                         exact=True)
    compare_cuda_and_cvm([x, i], [pt.set_subtensor(x[1:, i[:3]], 2.0), pt.inc_subtensor(x[::2, :, i], -1.0)], [xv, iv], exact=True)
    # duplicates accumulate (np.add.at)
    dup = np.array([2, 2, 2, 5, 5])
    compare_cuda_and_cvm([x, i, j], [pt.inc_subtensor(x[i, 2:4, j], 0.25)], [xv, dup, np.array([3, 3, 3, 1, 1])])
    # masked assignment of a vector of the selected size
    w = pt.fvector("w")
    sel = np.flatnonzero(mv).size
    if os.environ.get("PTK_DRY") != "1":
        compare_cuda_and_cvm([x, m3, w], [pt.set_subtensor(x[m3], w)], [xv, mv, rng.standard_normal(sel).astype("float32")],
                             exact=True)


def test_ignore_duplicates_increment(gpu):
    rng, xv = _data(44, (50,))
    x, i = pt.fvector("x"), pt.lvector("i")
    iv = rng.permutation(50)[:20]  # distinct: both semantics agree, the ignore_duplicates code path runs
    out = pt.inc_subtensor(x[i], 2.0, ignore_duplicates=True)
    compare_cuda_and_cvm([x, i], [out], [xv, iv], exact=True)


def test_out_of_bounds_and_mask_shape_errors(gpu):
    rng, xv = _data(45)
    x = pt.ftensor3("x")
    i, j = pt.lvector("i"), pt.lvector("j")
    f = pytensor.function([x, i, j], x[i, :, j], mode="CUDA")
    with pytest.raises(IndexError):
        f(xv, np.array([0, 7]), np.array([0, 0]))
    f(xv, np.array([0, 6]), np.array([0, 0]))
    m2 = pt.matrix("m2", dtype="bool")
    g = pytensor.function([x, m2], x[m2], mode="CUDA")
    with pytest.raises(IndexError):
        g(xv, np.ones((7, 8), dtype=bool))
