"""RandomVariable on the device (vm/nodes_random.py, csrc/ptk_random.cu) against the reference's contract
(pytensor/tensor/random/op.py:49, perform :457-468).  The device stream is counter-based Philox, not numpy's PCG64 with its
sequential rejection samplers, so VALUES differ from the C linker's by construction.  What must agree — and is tested — is:
shape / dtype / broadcasting of every draw (exact), the generator protocol (a call without `updates` repeats, with
`updates={rng: next_rng}` advances, equal seeds give equal streams), and the DISTRIBUTION of the draws: first two moments
within 5 standard errors and a Kolmogorov-Smirnov test (p > 1e-4) against scipy's cdf, n = 200000 per distribution."""

import numpy as np
import pytest
import scipy.stats as st

from helpers import pytensor

import pytensor.tensor as pt

pytestmark = pytest.mark.gpu

N = 200_000


def _draw(build, seed=123, n_calls=1, mode="CUDA"):
    rng = pytensor.shared(np.random.default_rng(seed), name="rng")
    nr, x = build(rng).owner.outputs
    f = pytensor.function([], x, updates={rng: nr}, mode=mode)
    return [np.asarray(f()) for _ in range(n_calls)], f


CASES = [
    ("uniform", lambda r: pt.random.uniform(-2.0, 3.0, size=(N,), rng=r), st.uniform(-2.0, 5.0)),
    ("normal", lambda r: pt.random.normal(1.5, 0.5, size=(N,), rng=r), st.norm(1.5, 0.5)),
    ("halfnormal", lambda r: pt.random.halfnormal(0.0, 2.0, size=(N,), rng=r), st.halfnorm(0.0, 2.0)),
    ("lognormal", lambda r: pt.random.lognormal(0.2, 0.4, size=(N,), rng=r), st.lognorm(0.4, scale=np.exp(0.2))),
    ("exponential", lambda r: pt.random.exponential(2.5, size=(N,), rng=r), st.expon(scale=2.5)),
    ("laplace", lambda r: pt.random.laplace(-1.0, 0.7, size=(N,), rng=r), st.laplace(-1.0, 0.7)),
    ("logistic", lambda r: pt.random.logistic(0.5, 1.2, size=(N,), rng=r), st.logistic(0.5, 1.2)),
    ("gumbel", lambda r: pt.random.gumbel(0.3, 1.1, size=(N,), rng=r), st.gumbel_r(0.3, 1.1)),
    ("cauchy", lambda r: pt.random.cauchy(0.0, 1.0, size=(N,), rng=r), st.cauchy(0.0, 1.0)),
    ("gamma_big", lambda r: pt.random.gamma(4.5, scale=0.5, size=(N,), rng=r), st.gamma(4.5, scale=0.5)),
    ("gamma_small", lambda r: pt.random.gamma(0.3, scale=2.0, size=(N,), rng=r), st.gamma(0.3, scale=2.0)),
    ("beta", lambda r: pt.random.beta(2.0, 3.5, size=(N,), rng=r), st.beta(2.0, 3.5)),
    ("weibull", lambda r: pt.random.weibull(1.7, size=(N,), rng=r), st.weibull_min(1.7)),
    ("invgamma", lambda r: pt.random.invgamma(5.0, 2.0, size=(N,), rng=r), st.invgamma(5.0, scale=2.0)),
    ("pareto", lambda r: pt.random.pareto(3.5, 2.0, size=(N,), rng=r), st.pareto(3.5, scale=2.0)),
    ("halfcauchy", lambda r: pt.random.halfcauchy(0.0, 1.5, size=(N,), rng=r), st.halfcauchy(0.0, 1.5)),
    ("studentt", lambda r: pt.random.t(7.0, 0.5, 1.5, size=(N,), rng=r), st.t(7.0, 0.5, 1.5)),
]


@pytest.mark.parametrize("name,build,dist", CASES, ids=[c[0] for c in CASES])
def test_continuous_distributions_moments_and_ks(gpu, name, build, dist):
    pytensor.config.floatX = "float64"
    (x,), _ = _draw(build)
    assert x.shape == (N,) and x.dtype == np.float64 and np.all(np.isfinite(x))
    ks = st.kstest(x, dist.cdf)
    assert ks.pvalue > 1e-4, (name, ks)
    if name not in ("cauchy", "halfcauchy"):
        m, v = dist.mean(), dist.var()
        kurt = float(dist.stats(moments="k")) + 3.0
        assert abs(x.mean() - m) < 5 * np.sqrt(v / N), (name, x.mean(), m)
        if np.isfinite(kurt):   # (the variance of a sample variance needs the fourth moment: Pareto(3.5) has none)
            assert abs(x.var() - v) < 5 * v * np.sqrt(max(kurt - 1.0, 2.0) / N) + 1e-12, (name, x.var(), v)


def test_discrete_distributions(gpu):
    pytensor.config.floatX = "float64"
    (b,), _ = _draw(lambda r: pt.random.bernoulli(0.3, size=(N,), rng=r))
    assert b.dtype.kind in "iub" and set(np.unique(b)) <= {0, 1} and abs(b.mean() - 0.3) < 5 * np.sqrt(0.21 / N)
    (k,), _ = _draw(lambda r: pt.random.integers(-3, 9, size=(N,), rng=r))
    assert k.dtype == np.int64 and k.min() == -3 and k.max() == 8
    counts = np.bincount(k + 3, minlength=12)
    assert st.chisquare(counts).pvalue > 1e-4


def test_generator_protocol_shapes_dtypes_and_broadcasting(gpu):
    pytensor.config.floatX = "float32"
    # shape / dtype agree with the reference for size=None broadcasting and for an explicit size
    mu = pt.fmatrix("mu")
    sd = pt.fvector("sd")
    rng = pytensor.shared(np.random.default_rng(7))
    outs = [pt.random.normal(mu, sd, rng=rng), pt.random.normal(mu, sd, size=(4, 3, 5), rng=rng),
            pt.random.uniform(0, 1, size=(2, 0), rng=rng)]
    muv = np.arange(15, dtype="float32").reshape(3, 5) * 100
    sdv = np.full(5, 1e-3, dtype="float32")
    got = pytensor.function([mu, sd], outs, mode="CUDA")(muv, sdv)
    ref = pytensor.function([mu, sd], outs, mode="CVM")(muv, sdv)
    for g, e in zip(got, ref):
        assert np.asarray(g).shape == np.asarray(e).shape and np.asarray(g).dtype == np.asarray(e).dtype
    np.testing.assert_allclose(got[0], muv, atol=0.01)                       # parameters land on the right elements
    np.testing.assert_allclose(got[1], np.broadcast_to(muv, (4, 3, 5)), atol=0.01)
    # no update: the same draws again; with the update: a new block each call; equal seeds: equal streams
    x = pt.random.normal(0, 1, size=(1000,), rng=rng)
    f_same = pytensor.function([], x, mode="CUDA")
    np.testing.assert_array_equal(f_same(), f_same())
    (a1, a2), _ = _draw(lambda r: pt.random.normal(0, 1, size=(1000,), rng=r), seed=5, n_calls=2)
    (b1, b2), _ = _draw(lambda r: pt.random.normal(0, 1, size=(1000,), rng=r), seed=5, n_calls=2)
    assert not np.array_equal(a1, a2)
    np.testing.assert_array_equal(a1, b1)
    np.testing.assert_array_equal(a2, b2)
    assert abs(np.corrcoef(a1, a2)[0, 1]) < 0.15


def test_draws_feed_device_kernels(gpu):
    # a Monte-Carlo estimate inside one compiled function: E[exp(-z^2)] for z ~ N(0,1) is 1/sqrt(3)
    pytensor.config.floatX = "float64"
    rng = pytensor.shared(np.random.default_rng(11))
    nr, z = pt.random.normal(0, 1, size=(400_000,), rng=rng).owner.outputs
    f = pytensor.function([], pt.exp(-z * z).mean(), updates={rng: nr}, mode="CUDA")
    est = np.mean([f() for _ in range(3)])
    assert abs(est - 1 / np.sqrt(3)) < 2e-3
