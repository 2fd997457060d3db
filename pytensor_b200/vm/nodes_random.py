"""RandomVariable on the device (reference: pytensor/tensor/random/op.py:49; perform :457-468 = `rng_fn(rng, *params, size)`
on a host numpy Generator, returning (the advanced generator, the draws)).

The generator stays a HOST object: each call takes 128 bits from it (which advances it exactly as a draw would, so the
`updates={rng: next_rng}` contract and copy-vs-inplace semantics are the reference's), and those bits key a counter-based
Philox stream on the device (csrc/ptk_random.cu).  Shapes, dtypes and parameter broadcasting follow the reference; the
VALUES are a different (equally distributed) stream — parity is distributional, see tests/test_gpu_random.py.
A graph with RandomVariable nodes is never captured into a CUDA graph (the key changes on every call)."""

from __future__ import annotations

import copy

import numpy as np

from ..runtime import device as dev
from ..runtime import lib as _lib
from .nodes_basic import _broadcast_view
from .nodes_elemwise import Node
from .values import Val

# reference op name -> (ptk_random_fill code, number of distribution parameters it takes)
DIST = {"uniform": (0, 2), "normal": (1, 2), "halfnormal": (2, 2), "lognormal": (3, 2), "exponential": (4, 1), "laplace": (5, 2),
        "logistic": (6, 2), "gumbel": (7, 2), "cauchy": (8, 2), "bernoulli": (9, 1), "gamma": (10, 2), "beta": (11, 2),
        "integers": (12, 2), "weibull": (13, 1), "pareto": (14, 2), "halfcauchy": (15, 2), "invgamma": (16, 2),
        "studentt": (17, 3), "t": (17, 3)}


def _generator(v: Val):
    g = v.h
    if isinstance(g, np.ndarray):
        g = g.item()
    if not isinstance(g, np.random.Generator):
        raise TypeError(f"RandomVariable: expected a numpy Generator, got {type(g).__name__}")
    return g


class RandomVariableNode(Node):
    def __init__(self, dist_name, dtype, inplace, size_is_none, name="RandomVariable"):
        self.code, self.n_params = DIST[dist_name]
        self.dist_name, self.dtype, self.inplace, self.size_is_none, self.name = dist_name, dtype, inplace, size_is_none, name

    def run(self, vals):
        if dev.alloc_state.capturing:
            raise dev.GraphUnsupported("random draws are keyed per call")
        gen = _generator(vals[0])
        if not self.inplace:
            gen = copy.deepcopy(gen)
        key, seed = (int(w) for w in gen.bit_generator.random_raw(2))   # advances the generator: the next call differs
        params = vals[2:2 + self.n_params]
        pshapes = [tuple(p.shape) for p in params]
        if self.size_is_none:
            shape = tuple(np.broadcast_shapes(*pshapes)) if pshapes else ()
        else:
            shape = tuple(int(s) for s in np.asarray(vals[1].host()).reshape(-1))
            if pshapes:
                np.broadcast_shapes(shape, *pshapes)   # raises like numpy when the parameters do not fit `size`
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        out = dev.empty(shape, self.dtype)
        ptrs, strides, keep = [], [], []
        for p in params:
            t = p.dev()
            if dev.TORCH_TO_NP[t.dtype] != "float64":
                from .nodes_cast import cast_to

                t = cast_to(t, "float64")
            if t.numel() == 1:
                ptrs.append(dev.ptr(t))
                strides.append(0)
            else:
                tb = _broadcast_view(t, shape)
                tb = tb if tb.is_contiguous() else dev.contiguous(tb)
                ptrs.append(dev.ptr(tb))
                strides.append(1)
                keep.append(tb)
            keep.append(t)
        while len(ptrs) < 3:
            ptrs.append(None)
            strides.append(0)
        if n:
            _lib.check(_lib.lib().ptk_random_fill(self.code, _lib.DTYPE_CODE[self.dtype], dev.ptr(out), n, key, seed, ptrs[0],
                                                  strides[0], ptrs[1], strides[1], ptrs[2], strides[2], dev.stream_ptr()),
                       "ptk_random_fill")
        return [Val(h=gen), Val(d=out)]
