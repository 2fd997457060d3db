"""Parity of the fused Elemwise kernels (K1) against the reference C linker.  Shapes follow the reference's
TestBroadcast sweep (tests/tensor/test_elemwise.py:259-302) plus the cfg-2 style Composite."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu

BROADCAST_SHAPES = [
    ((5, 6), (5, 6)), ((5, 6), (1, 6)), ((5, 6), (1, 1)), ((1, 6), (5, 6)), ((1, 1), (5, 6)),
    ((2, 3, 4, 5), (2, 3, 4, 5)), ((2, 3, 4, 5), (1, 3, 1, 5)), ((2, 3, 4, 5), (1, 1, 1, 1)), ((), ()),
    ((100, 64), (100, 64)), ((64, 128), (64, 1)), ((257, 33), (1, 33)), ((0, 6), (0, 6)), ((5, 0), (1, 0)),
]


@pytest.mark.parametrize("xsh,ysh", BROADCAST_SHAPES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_broadcast_add_mul(gpu, xsh, ysh, dtype):
    rng = np.random.default_rng(11)
    x = pt.tensor("x", dtype=dtype, shape=tuple(1 if s == 1 else None for s in xsh))
    y = pt.tensor("y", dtype=dtype, shape=tuple(1 if s == 1 else None for s in ysh))
    xv = rng.standard_normal(xsh).astype(dtype)
    yv = rng.standard_normal(ysh).astype(dtype)
    compare_cuda_and_cvm([x, y], [x + y, x * y - y, pt.exp(x) / (1 + pt.sqr(y))], [xv, yv])


def test_cfg2_style_composite(gpu):
    pytensor.config.floatX = "float32"
    rng = np.random.default_rng(1)
    a = pt.fmatrix("a")
    b = pt.fmatrix("b")
    e = a
    for c in [0.5, -0.25, 0.125, 0.75]:
        e = (e * b + np.float32(c)) * np.float32(0.9)
        e = pt.maximum(e, -e) + pt.sqr(a) * np.float32(0.1)
    e = pt.tanh(e * np.float32(0.01)) + pt.exp(-pt.abs(b))
    r = e.sum(axis=1)
    av = rng.standard_normal((515, 1028)).astype("float32")
    bv = rng.standard_normal((515, 1028)).astype("float32")
    f, _ = compare_cuda_and_cvm([a, b], [e, r], [av, bv])
    # odd sizes: tails, unaligned rows
    av = rng.standard_normal((37, 1001)).astype("float32")
    bv = rng.standard_normal((37, 1001)).astype("float32")
    compare_cuda_and_cvm([a, b], [e, r], [av, bv])


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 1024, 1025, 70001])
def test_flat_tails(gpu, n):
    rng = np.random.default_rng(n)
    x = pt.fvector("x")
    xv = rng.standard_normal(n).astype("float32")
    compare_cuda_and_cvm([x], [pt.tanh(x) * 2 + x, pt.switch(x > 0, x, pt.expm1(x))], [xv])


def test_transposed_and_strided_inputs(gpu):
    rng = np.random.default_rng(5)
    x = pt.dmatrix("x")
    y = pt.dmatrix("y")
    xv = rng.standard_normal((33, 65))
    yv = rng.standard_normal((65, 33))
    compare_cuda_and_cvm([x, y], [x + y.T, (x.T * y)[::2, 1::3], pt.exp(x[:, ::-1]) + x], [xv, yv])


SCALAR_FNS = [
    "exp", "log", "log2", "log10", "log1p", "expm1", "sqrt", "sin", "cos", "tan", "arcsin", "arccos", "arctan",
    "sinh", "cosh", "tanh", "arcsinh", "arctanh", "erf", "erfc", "gamma", "gammaln", "sigmoid", "softplus",
    "log1mexp", "abs", "sign", "ceil", "floor", "trunc", "round", "neg", "sqr", "reciprocal",
]


@pytest.mark.parametrize("fn", SCALAR_FNS)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_scalar_op_table(gpu, fn, dtype):
    rng = np.random.default_rng(3)
    x = pt.tensor("x", dtype=dtype, shape=(None,))
    xv = rng.uniform(0.05, 0.95, size=777).astype(dtype)
    if fn in ("exp", "tanh", "sinh", "cosh", "erf", "erfc", "sigmoid", "softplus", "abs", "sign", "ceil", "floor",
              "trunc", "round", "neg", "sqr", "arcsinh", "arctan", "sin", "cos", "tan", "expm1"):
        xv = (xv * 8 - 4).astype(dtype)
    if fn == "log1mexp":
        xv = (-xv * 5).astype(dtype)
    if fn in ("gamma", "gammaln"):
        xv = (xv * 6 + 0.1).astype(dtype)
    op = getattr(pt, fn) if hasattr(pt, fn) else getattr(pt.math, fn)
    compare_cuda_and_cvm([x], [op(x)], [xv], rtol=2e-5 if dtype == "float32" else 1e-10)


def test_integer_and_bool_semantics(gpu):
    rng = np.random.default_rng(9)
    i = pt.lvector("i")
    j = pt.lvector("j")
    k = pt.bvector("k")
    iv = rng.integers(-50, 50, size=301)
    jv = rng.integers(1, 9, size=301) * rng.choice([-1, 1], size=301)
    kv = rng.integers(-100, 100, size=301).astype("int8")
    outs = [i // j, i % j, i / j, i + k, pt.maximum(i, j), pt.abs(k), pt.eq(i, j), (i > j) & (k > 0), i ** 2,
            pt.cast(i, "float32") * 0.5, pt.switch(i > 0, i, j), pt.sign(i), ~(i > j), i & j, i | j, i ^ j]
    compare_cuda_and_cvm([i, j, k], outs, [iv, jv, kv])


def test_nan_propagating_maximum(gpu):
    x = pt.dvector("x")
    y = pt.dvector("y")
    xv = np.array([1.0, np.nan, 3.0, -np.inf, np.nan])
    yv = np.array([2.0, 1.0, np.nan, np.inf, np.nan])
    compare_cuda_and_cvm([x, y], [pt.maximum(x, y), pt.minimum(x, y), pt.isnan(x), pt.isinf(y)], [xv, yv])


def test_inplace_and_multi_output(gpu):
    rng = np.random.default_rng(2)
    x = pt.dmatrix("x")
    y = pt.dvector("y")
    z = pt.exp(x) + y
    outs = [z * 2, z - x, pt.tanh(z)]
    compare_cuda_and_cvm([x, y], outs, [rng.standard_normal((40, 24)), rng.standard_normal(24)])


def test_runtime_broadcast_error(gpu):
    # the reference forbids broadcasting a dim that is not typed broadcastable (tests/tensor/test_elemwise.py:839)
    x = pt.dmatrix("x")
    y = pt.dmatrix("y")
    f = pytensor.function([x, y], x + y, mode="CUDA")
    with pytest.raises(ValueError):
        f(np.ones((3, 1)), np.ones((3, 4)))
