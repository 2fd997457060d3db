"""Special functions and the softmax family on the device against the reference C linker.

Psi / TriGamma / GammaInc / GammaIncC / BetaInc are the scalar ops whose C implementations the reference ships as
support code (pytensor/scalar/math.py:403,502,627,674,1342; scalar/c_code/gamma.c, incbet.c) and which the log-densities
of Gamma / Beta / StudentT-type models and their gradients need.  Softmax / LogSoftmax / logsumexp are symbolic in this
reference version (pytensor/tensor/special.py) and must lower to the Elemwise / CAReduce kernels."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_psi_trigamma(gpu, dtype):
    pytensor.config.floatX = dtype
    rng = np.random.default_rng(31)
    x = pt.matrix("x", dtype=dtype)
    xv = rng.uniform(-20, 50, (64, 129)).astype(dtype)
    xv[0, :12] = [0.0, -1.0, -2.0, -0.5, 1e-6, 1e-5, 9e-5, 1e-4, 8.5, 5.0, 1e4, -1e-3]
    compare_cuda_and_cvm([x], [pt.psi(x), pt.tri_gamma(x), pt.gammaln(pt.abs(x) + 0.5) - pt.psi(pt.abs(x) + 0.5)], [xv])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_gammainc_gammaincc_regimes_and_limits(gpu, dtype):
    # reference tests: tests/scalar/test_math.py:26-80 (known values, nan and inf arguments)
    pytensor.config.floatX = dtype
    rng = np.random.default_rng(32)
    k, x = pt.vector("k", dtype=dtype), pt.vector("x", dtype=dtype)
    kv = np.concatenate([rng.uniform(0.05, 40, 5000), [1, 1, np.inf, 1, np.inf, -1, 1, 0.5, 170, 171.5, 300]]).astype(dtype)
    xv = np.concatenate([rng.uniform(0, 80, 5000), [2, np.inf, 1, np.nan, np.inf, 1, -1, 0, 160, 180, 310]]).astype(dtype)
    compare_cuda_and_cvm([k, x], [pt.gammainc(k, x), pt.gammaincc(k, x)], [kv, xv])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_betainc_regimes_and_limits(gpu, dtype):
    pytensor.config.floatX = dtype
    rng = np.random.default_rng(33)
    a, b, x = (pt.vector(n, dtype=dtype) for n in "abx")
    av = np.concatenate([rng.uniform(0.05, 30, 5000), [1, 2, 0.5, 100, 200, 1e-2, 5, -1, 3, 3]]).astype(dtype)
    bv = np.concatenate([rng.uniform(0.05, 30, 5000), [1, 3, 0.5, 100, 150, 50, 1e-2, 2, -2, 3]]).astype(dtype)
    xv = np.concatenate([rng.uniform(0, 1, 5000), [0.3, 0.0, 1.0, 0.5, 0.6, 0.01, 0.99, 0.5, 0.5, 1.5]]).astype(dtype)
    compare_cuda_and_cvm([a, b, x], [pt.betainc(a, b, x)], [av, bv, xv])


def test_gamma_and_beta_logp_with_gradients(gpu):
    """PyMC-style Gamma(alpha, beta) and Beta(a, b) log-densities summed over data, with gradients w.r.t. the
    parameters: the gradient graphs contain Psi (d lgamma) next to the usual Elemwise / CAReduce nodes."""
    pytensor.config.floatX = "float64"
    rng = np.random.default_rng(34)
    alpha, beta, a, b = pt.dscalars("alpha", "beta", "a", "b")
    y, u = pt.dvector("y"), pt.dvector("u")
    logp_gamma = (alpha * pt.log(beta) - pt.gammaln(alpha) + (alpha - 1) * pt.log(y) - beta * y).sum()
    logp_beta = (pt.gammaln(a + b) - pt.gammaln(a) - pt.gammaln(b) + (a - 1) * pt.log(u) + (b - 1) * pt.log1p(-u)).sum()
    logp = logp_gamma + logp_beta
    grads = pytensor.grad(logp, [alpha, beta, a, b])
    yv, uv = rng.gamma(2.0, 1.5, 4097), rng.beta(2.0, 3.0, 4097)
    compare_cuda_and_cvm([alpha, beta, a, b, y, u], [logp, *grads], [2.5, 1.25, 1.75, 3.5, yv, uv])


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("axis", [-1, 0, None])
def test_softmax_logsoftmax_logsumexp(gpu, dtype, axis):
    pytensor.config.floatX = dtype
    rng = np.random.default_rng(35)
    x = pt.matrix("x", dtype=dtype)
    xv = (rng.standard_normal((67, 301)) * 4).astype(dtype)
    from pytensor.tensor.special import log_softmax, softmax

    outs = [softmax(x, axis=axis), log_softmax(x, axis=axis), pt.logsumexp(x, axis=axis)]
    compare_cuda_and_cvm([x], outs, [xv])


def test_softmax_gradient(gpu):
    pytensor.config.floatX = "float32"
    rng = np.random.default_rng(36)
    x, t = pt.fmatrix("x"), pt.fmatrix("t")
    from pytensor.tensor.special import log_softmax

    loss = -(t * log_softmax(x, axis=-1)).sum(axis=-1).mean()
    g = pytensor.grad(loss, x)
    xv = rng.standard_normal((128, 1000)).astype("float32")
    tv = np.eye(1000, dtype="float32")[rng.integers(0, 1000, 128)]
    compare_cuda_and_cvm([x, t], [loss, g], [xv, tv])
