"""Parity of CAReduce kernels (K2) — cases follow the reference's TestCAReduce sweep (tests/tensor/test_elemwise.py
:446-462 shapes/axes incl. empty dims, :598-609 ops/dtypes, :611-638 NaN propagation)."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu

CASES = [
    ((5, 6), None), ((5, 6), (0, 1)), ((5, 6), (0,)), ((5, 6), (1,)), ((5, 6), (-1,)), ((5, 6), ()),
    ((2, 3, 4, 5), (0, 1, 3)), ((2, 3, 4, 5), (-2, -3)), ((2, 3, 4, 5), (1, 2)), ((2, 3, 4, 5), (3,)),
    ((2, 3, 4, 5), (0, 2)), ((5, 0), None), ((5, 0), (0,)), ((5, 0), (1,)), ((0, 0), None), ((), None), ((), ()),
    ((300, 257), (1,)), ((300, 257), (0,)), ((3, 70000), (1,)), ((70000, 3), (0,)), ((1 << 18,), None),
    ((64, 32, 48), (1,)), ((64, 32, 48), (0, 2)),
]


@pytest.mark.parametrize("shape,axis", CASES)
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_sum_max_min_prod(gpu, shape, axis, dtype):
    rng = np.random.default_rng(17)
    x = pt.tensor("x", dtype=dtype, shape=(None,) * len(shape))
    xv = rng.uniform(0.5, 1.5, size=shape).astype(dtype)
    outs = [pt.sum(x, axis=axis)]
    if int(np.prod(shape)) > 0 or axis == ():
        outs += [pt.max(x, axis=axis), pt.min(x, axis=axis)]
    if int(np.prod(shape)) < 5000:
        outs.append(pt.prod(x, axis=axis))
    compare_cuda_and_cvm([x], outs, [xv])


@pytest.mark.parametrize("dtype", ["int8", "uint8", "int32", "int64", "bool"])
def test_integer_reductions(gpu, dtype):
    rng = np.random.default_rng(4)
    x = pt.tensor("x", dtype=dtype, shape=(None, None))
    xv = rng.integers(0, 2 if dtype == "bool" else 100, size=(37, 53)).astype(dtype)
    outs = [pt.sum(x, axis=1), pt.sum(x), pt.all(x, axis=0), pt.any(x, axis=1)]
    if dtype != "bool":
        outs += [pt.max(x, axis=0), pt.min(x, axis=1)]
    compare_cuda_and_cvm([x], outs, [xv])


def test_nan_propagation(gpu):
    x = pt.dmatrix("x")
    xv = np.arange(30, dtype="float64").reshape(5, 6)
    xv[2, 3] = np.nan
    compare_cuda_and_cvm([x], [pt.max(x, axis=1), pt.min(x, axis=0), pt.sum(x, axis=0), pt.max(x)], [xv])


def test_fp32_sum_accumulates_in_fp64(gpu):
    # Sum{acc=float64} (pytensor/tensor/elemwise.py:1383-1397): a long fp32 sum must match the C linker to 1e-6
    rng = np.random.default_rng(8)
    x = pt.fmatrix("x")
    xv = (rng.standard_normal((16, 200000)) * 100).astype("float32")
    compare_cuda_and_cvm([x], [x.sum(axis=1), x.sum()], [xv], rtol=1e-6)


def test_noncontiguous_input(gpu):
    rng = np.random.default_rng(6)
    x = pt.dtensor3("x")
    xv = rng.standard_normal((7, 9, 11))
    compare_cuda_and_cvm([x], [x.transpose(2, 0, 1).sum(axis=1), x[:, ::2, :].max(axis=(0, 2))], [xv])
