"""More of the Op library at thunk level: ARange, Eye, ExtractDiag, Split, Argmax, CumOp against the
reference C linker (index results bit-exact).  Edge cases follow the reference's own tests: ties / NaNs for argmax
(tests/tensor/test_math.py TestMaxAndArgmax), zero-length pieces for split (tests/tensor/test_basic.py TestJoinAndSplit),
offsets and non-square shapes for eye / diagonal, every axis for cumsum / cumprod (tests/tensor/test_extra_ops.py)."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,axis", [((5, 7), 0), ((5, 7), 1), ((5, 7), None), ((300, 5000), 1), ((4099, 3), 0),
                                         ((70, 3000), 0), ((3, 4, 5), (0, 2)), ((6, 5, 4), (1, 2)), ((1, 1), None)])
@pytest.mark.parametrize("dtype", ["float32", "int32"])
def test_argmax_first_maximum_ties_and_nans(gpu, shape, axis, dtype):
    rng = np.random.default_rng(91)
    T = pt.TensorType(dtype, shape=(None,) * len(shape))
    x = T("x")
    if dtype == "float32":
        xv = rng.integers(-3, 4, size=shape).astype(dtype)  # many ties: the FIRST maximum must win
        if xv.size > 20:
            flat = xv.reshape(-1)
            flat[rng.integers(0, flat.size, size=max(1, flat.size // 50))] = np.nan  # a NaN is the maximum (np.argmax)
    else:
        xv = rng.integers(-100, 100, size=shape).astype(dtype)
    compare_cuda_and_cvm([x], [pt.argmax(x, axis=axis)], [xv], exact=True)


def test_max_and_argmax_float64(gpu):
    x = pt.dmatrix("x")
    xv = np.random.default_rng(92).standard_normal((33, 65))
    mx, am = pt.max_and_argmax(x, axis=1)
    compare_cuda_and_cvm([x], [mx, am], [xv])


def test_split_views_zero_length_piece_and_errors(gpu):
    x = pt.dmatrix("x")
    s = pt.lvector("s")
    a, b, c = pt.split(x, s, n_splits=3, axis=1)
    xv = np.random.default_rng(93).standard_normal((6, 9))
    compare_cuda_and_cvm([x, s], [a * 2, b, c.sum(axis=1)], [xv, np.array([4, 0, 5])])
    f = pytensor.function([x, s], [a, c], mode="CUDA")
    if gpu:
        with pytest.raises(ValueError):
            f(xv, np.array([4, 1, 5]))  # sizes do not sum to the axis length
        with pytest.raises(ValueError):
            f(xv, np.array([10, -1, 0]))


@pytest.mark.parametrize("dtype", ["int64", "float32", "float64", "int32"])
def test_arange_host_and_device_sizes(gpu, dtype):
    n = pt.lscalar("n")
    x = pt.fvector("x")
    # short (host-visible index plumbing) and long (ptk_arange) ranges; start/step in the output type
    start, step = (np.asarray(3, dtype), np.asarray(2, dtype)) if "int" in dtype else (np.asarray(0.5, dtype), np.asarray(0.25, dtype))
    out = pt.arange(start, start + step * n.astype(dtype), step, dtype=dtype)
    for nv in (7, 100_003):
        f, got = compare_cuda_and_cvm([n], [out, out[::-1] * 2], [np.asarray(nv, "int64")], exact=("int" in dtype), rtol=1e-6)
    # from a shape, feeding an elementwise op (the triu / tri construction)
    compare_cuda_and_cvm([x], [pt.arange(x.shape[0]) * 2 + 1, pt.triu(pt.outer(x, x)), pt.tril(pt.outer(x, x), k=-1)],
                         [np.random.default_rng(94).standard_normal(37).astype("float32")])


@pytest.mark.parametrize("n,m,k", [(5, 5, 0), (4, 7, 2), (7, 4, -3), (3, 3, 5), (6, 2, -1), (0, 3, 0)])
def test_eye(gpu, n, m, k):
    a, b = pt.lscalar("a"), pt.lscalar("b")
    for dtype in ("float32", "int64"):
        compare_cuda_and_cvm([a, b], [pt.eye(a, b, k, dtype=dtype)], [np.asarray(n, "int64"), np.asarray(m, "int64")],
                             exact=True)


def test_extract_diag_trace_and_offsets(gpu):
    x = pt.dmatrix("x")
    t = pt.dtensor3("t")
    rng = np.random.default_rng(95)
    xv, tv = rng.standard_normal((6, 9)), rng.standard_normal((4, 5, 6))
    compare_cuda_and_cvm([x], [pt.diag(x) * 2, pt.trace(x), pt.diagonal(x, offset=3), pt.diagonal(x, offset=-2) + 1], [xv])
    compare_cuda_and_cvm([t], [pt.diagonal(t, offset=1, axis1=0, axis2=2), pt.diagonal(t, axis1=1, axis2=2).sum(axis=0)],
                         [tv])


@pytest.mark.parametrize("shape,axis", [((7, 33), 0), ((7, 33), 1), ((129, 1000), 1), ((129, 1000), 0), ((100_001,), 0),
                                         ((5, 6, 7), 1), ((5, 6, 7), None)])
@pytest.mark.parametrize("dtype", ["float32", "float64", "int64"])
def test_cumsum_cumprod(gpu, shape, axis, dtype):
    rng = np.random.default_rng(96)
    T = pt.TensorType(dtype, shape=(None,) * len(shape))
    x = T("x")
    if dtype == "int64":
        xs = rng.integers(-5, 6, size=shape).astype(dtype)
        xp = rng.integers(-2, 3, size=shape).astype(dtype)  # products wrap modulo 2^64 exactly like NumPy's
    else:
        xs = rng.standard_normal(shape).astype(dtype)
        xp = (1.0 + 0.01 * rng.standard_normal(shape)).astype(dtype)  # products stay O(1)
    # last-axis scans use a warp tree + carried total instead of NumPy's strictly sequential loop: same values up to fp
    # reassociation, i.e. an absolute error relative to the magnitude of the running sums (1e-5 of it in fp32)
    scale = max(1.0, float(np.abs(np.cumsum(xs.astype(np.float64), axis=axis)).max()))
    tol = dict(rtol=1e-5, atol=1e-5 * scale) if dtype == "float32" else dict(rtol=1e-10, atol=1e-12 * scale)
    compare_cuda_and_cvm([x], [pt.cumsum(x, axis=axis)], [xs], exact=(dtype == "int64"), **tol)
    compare_cuda_and_cvm([x], [pt.cumprod(x, axis=axis)], [xp], exact=(dtype == "int64"), **tol)


def test_cumsum_gradient_graph(gpu):
    # the pullback of CumOp reverses, cumsums and reverses again (extra_ops.py:323-336): Subtensor views feeding CumOp
    x = pt.dmatrix("x")
    g = pytensor.grad((pt.cumsum(x, axis=1) ** 2).sum(), x)
    compare_cuda_and_cvm([x], [g], [np.random.default_rng(97).standard_normal((9, 17))])


def test_advanced_indexing_with_several_index_arrays(gpu):
    # NumPy advanced indexing with k index arrays on consecutive axes (tests/tensor/test_subtensor.py TestAdvancedSubtensor):
    # broadcasting between the index arrays, negative indices, leading / trailing full axes, gather + set + inc
    rng = np.random.default_rng(98)
    x = pt.dtensor3("x")
    i, j = pt.lvector("i"), pt.lmatrix("j")
    xv = rng.standard_normal((5, 6, 7))
    iv = np.array([0, -1, 3, 3], dtype="int64")
    jv = rng.integers(-6, 6, size=(3, 4)).astype("int64")
    y = pt.dmatrix("y")
    yv = rng.standard_normal((3, 4))
    outs = [x[i, i], x[i, j], x[:, i, j], x[:, :, i][:, j[0]], x[i, j, i]]
    compare_cuda_and_cvm([x, i, j], outs, [xv, iv, jv], exact=True)
    outs = [pt.set_subtensor(x[i, j], 1.5), pt.inc_subtensor(x[:, i, j], y), pt.inc_subtensor(x[i, j, i], 2.0)]
    compare_cuda_and_cvm([x, i, j, y], outs, [xv, iv, jv, yv])
    # duplicates accumulate (np.add.at): every pair is (0, 0)
    z = pt.dmatrix("z")
    compare_cuda_and_cvm([z], [pt.inc_subtensor(z[[0, 0, 0], [0, 0, 0]], 1.0)], [np.zeros((2, 2))])
    # pt.diag(vector) = AllocDiag: zeros + arange + two-array set (tensor/basic.py:3903-3913)
    v = pt.dvector("v")
    compare_cuda_and_cvm([v], [pt.diag(v), pt.diag(v, k=2), pt.diag(v, k=-1) @ pt.diag(v, k=1)], [rng.standard_normal(6)])
    f = pytensor.function([x, i, j], x[i, j], mode="CUDA")
    if gpu:
        with pytest.raises(IndexError):
            f(xv, np.array([0, 5, 0, 0], dtype="int64"), jv)  # 5 is out of range for axis 0 (length 5)
