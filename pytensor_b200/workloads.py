"""PyTensor graph builders for the BASELINE.json configs (SURVEY.md §8d gives the concrete synthetic recipes).

Pure graph construction through the host's public API — used by bench.py, the tests and `__graft_entry__.smoke()`.
Every builder returns (inputs, outputs, make_args(rng) -> list of NumPy arrays, meta).
"""

from __future__ import annotations

import numpy as np


def _pt():
    from pytensor_b200._host import ensure_pytensor

    ensure_pytensor()
    import pytensor
    import pytensor.tensor as pt

    return pytensor, pt


def cfg1_readme(n=1024):
    """configs[0]: README `a/a + (M+a).dot(v)`, fp64 (README.rst:38-61)."""
    pytensor, pt = _pt()
    a, v, M = pt.dscalar("a"), pt.dvector("v"), pt.dmatrix("M")
    out = a / a + (M + a).dot(v)

    def make_args(seed=0):
        rng = np.random.default_rng(seed)
        return [np.float64(1.5), np.ones(n), rng.standard_normal((n, n))]

    return [a, v, M], [out], make_args, {"bytes": 8 * (n * n + 2 * n), "name": "cfg1_readme"}


def cfg2_fused_elemwise(n=4096, rounds=4):
    """configs[1]: 32-op fused Elemwise + CAReduce over fp32 (n,n): HBM-bound (<= 2 transcendentals per element).
    Outputs e (n,n) f32 and r = e.sum(axis=1) (acc f64)."""
    pytensor, pt = _pt()
    a, b = pt.fmatrix("a"), pt.fmatrix("b")
    f32 = np.float32
    e = a
    cs = [0.5, -0.25, 0.125, 0.75, -0.375, 0.0625]
    for c in cs[:rounds]:
        e = (e * b + f32(c)) * f32(0.9)
        e = pt.maximum(e, -e) + pt.sqr(a) * f32(0.1)
    e = pt.tanh(e * f32(0.01)) + pt.exp(-pt.abs(b))
    r = e.sum(axis=1)

    def make_args(seed=1):
        return [np.random.default_rng(seed).standard_normal((n, n)).astype("float32"),
                np.random.default_rng(seed + 1).standard_normal((n, n)).astype("float32")]

    # algorithmic bytes: read a, b; write e; write r  (201,342,976 B at n = 4096)
    return [a, b], [e, r], make_args, {"bytes": 3 * 4 * n * n + 4 * n, "name": "cfg2_fused_elemwise"}


def cfg3_mlp(n=4096, layers=3):
    """configs[2]: 3-layer MLP h = tanh(h @ W_i + b_i), fp32 at the graph boundary (no bf16 dtype in the host)."""
    pytensor, pt = _pt()
    x = pt.fmatrix("x")
    Ws = [pt.fmatrix(f"W{i}") for i in range(layers)]
    bs = [pt.fvector(f"b{i}") for i in range(layers)]
    h = x
    for W, b in zip(Ws, bs):
        h = pt.tanh(pt.dot(h, W) + b)

    def make_args(seed=3):
        args = [np.random.default_rng(seed).standard_normal((n, n)).astype("float32") / 64]
        for i in range(layers):
            args.append((np.random.default_rng(seed + 1 + i).standard_normal((n, n)) / 64).astype("float32"))
        for i in range(layers):
            args.append((np.random.default_rng(seed + 10 + i).standard_normal(n) / 64).astype("float32"))
        return args

    return [x, *Ws, *bs], [h], make_args, {"flops": layers * 2 * n ** 3, "name": "cfg3_mlp"}


def cfg4_scan(rows=8192, cols=512, n_steps=1000, matmul=False, full_trace=False):
    """configs[3]: Scan recurrence h <- tanh(h*a + b) (or tanh(h@W + b)), carried state (rows, cols) fp32."""
    pytensor, pt = _pt()
    h0 = pt.fmatrix("h0")
    b = pt.fvector("b")
    if matmul:
        W = pt.fmatrix("W")
        hs = pytensor.scan(lambda h, W, b: pt.tanh(pt.dot(h, W) + b), outputs_info=[h0], non_sequences=[W, b],
                           n_steps=n_steps, return_updates=False)
        ins = [h0, W, b]
    else:
        a = pt.fvector("a")
        hs = pytensor.scan(lambda h, a, b: pt.tanh(h * a + b), outputs_info=[h0], non_sequences=[a, b],
                           n_steps=n_steps, return_updates=False)
        ins = [h0, a, b]
    out = hs if full_trace else hs[-1]

    def make_args(seed=10):
        rng = np.random.default_rng(seed)
        args = [rng.standard_normal((rows, cols)).astype("float32")]
        if matmul:
            args.append((rng.standard_normal((cols, cols)) / np.sqrt(cols)).astype("float32"))
        else:
            args.append(rng.uniform(0.5, 1.5, cols).astype("float32"))
        args.append((rng.standard_normal(cols) * 0.1).astype("float32"))
        return args

    state = 4 * rows * cols
    return ins, [out], make_args, {"bytes": (n_steps if full_trace else 2) * state, "state_bytes": state,
                                   "n_steps": n_steps, "name": "cfg4_scan"}


def cfg5_logp_grad(B=1 << 20, n=1024, J=64, K=8, dtype="float32", packed=False):
    """configs[4]: hierarchical-normal logp + grad for B independent parameter vectors (chains) over shared data.

    logp_b = N(mu;0,1) + N(ls;0,1) + sum_j N(theta_j; mu, e^ls) + sum_k N(beta_k;0,1) + sum_i N(y_i; theta[g_i]+X_i.beta, 1)
    Outputs: [sum_b logp_b, d/dmu, d/dls, d/dtheta (J), d/dbeta (K)] each summed over the local batch — the vector
    that is all-reduced across GPUs.  Built directly in batched form (B leading), which is what `vectorize_graph`
    produces for this model."""
    pytensor, pt = _pt()
    T = lambda name, nd: pt.tensor(name, dtype=dtype, shape=(None,) * nd)  # noqa: E731
    mu, ls = T("mu", 1), T("ls", 1)            # (B,)
    theta, beta = T("theta", 2), T("beta", 2)  # (B,J), (B,K)
    y = T("y", 1)                              # (n,)
    X = T("X", 2)                              # (n,K)
    g = pt.lvector("g")                        # (n,)
    c = np.asarray(-0.5 * np.log(2 * np.pi), dtype=dtype)

    def normal_logp(x, m, log_s):
        z = (x - m) * pt.exp(-log_s)
        return c - log_s - 0.5 * z * z

    pred = theta[:, g] + pt.dot(beta, X.T)                       # (B,n)
    lp = (normal_logp(mu, 0.0, 0.0) + normal_logp(ls, 0.0, 0.0)
          + normal_logp(theta, mu[:, None], ls[:, None]).sum(axis=1)
          + normal_logp(beta, 0.0, 0.0).sum(axis=1)
          + normal_logp(y[None, :], pred, 0.0).sum(axis=1))      # (B,)
    total = lp.sum()
    grads = pytensor.grad(total, [mu, ls, theta, beta])
    outs = [total] + [gr.sum(axis=0) for gr in grads]
    if packed:  # ONE contiguous [logp, d/dmu, d/dls, d/dtheta (J), d/dbeta (K)] vector: the all-reduce message
        outs = [pt.concatenate([o.reshape((-1,)) for o in outs])]

    def make_args(seed=20, B_local=None):
        rng = np.random.default_rng(seed)
        Bl = B if B_local is None else B_local
        return [
            rng.standard_normal(Bl).astype(dtype) * 0.5, rng.standard_normal(Bl).astype(dtype) * 0.1,
            rng.standard_normal((Bl, J)).astype(dtype), rng.standard_normal((Bl, K)).astype(dtype) * 0.3,
            np.random.default_rng(seed + 1).standard_normal(n).astype(dtype),
            np.random.default_rng(seed + 2).standard_normal((n, K)).astype(dtype),
            np.random.default_rng(seed + 3).integers(0, J, size=n).astype("int64"),
        ]

    P = 2 + J + K
    return [mu, ls, theta, beta, y, X, g], outs, make_args, {
        "bytes": B * P * 4 + n * (K + 2) * 4 + (1 + P) * 4, "P": P, "name": "cfg5_logp_grad"}


def metric_graph(n=64, layers=84, scan_steps=16, dtype="float32"):
    """The 256-node class metric graph: 84 x tanh(h@W+b), a 16-step Scan, a final Sum.
    `dtype="float64"`: the same graph over the same (float32-valued) numbers in double — the bench's yardstick for how far
    two fp32 evaluations of this 84-layer chain may legitimately sit apart."""
    pytensor, pt = _pt()
    x = pt.matrix("x", dtype=dtype)
    Ws = [pt.matrix(f"W{i}", dtype=dtype) for i in range(layers)]
    bs = [pt.vector(f"b{i}", dtype=dtype) for i in range(layers)]
    a = pt.vector("a", dtype=dtype)
    h = x
    for W, b in zip(Ws, bs):
        h = pt.tanh(pt.dot(h, W) + b)
    c01 = np.asarray(np.float32(0.1), dtype=dtype)
    hs = pytensor.scan(lambda h, a: pt.tanh(h * a + c01), outputs_info=[h], non_sequences=[a],
                       n_steps=scan_steps, return_updates=False)
    out = hs[-1].sum(axis=0)

    def make_args(seed=40):
        rng = np.random.default_rng(seed)
        args = [rng.standard_normal((n, n)).astype("float32")]
        args += [(rng.standard_normal((n, n)) / np.sqrt(n)).astype("float32") for _ in range(layers)]
        args += [(rng.standard_normal(n) * 0.1).astype("float32") for _ in range(layers)]
        args += [rng.uniform(0.5, 1.5, n).astype("float32")]
        return [v.astype(dtype) for v in args]

    return [x, *Ws, *bs, a], [out], make_args, {"name": "metric_graph"}
