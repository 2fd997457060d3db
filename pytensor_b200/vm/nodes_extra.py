"""More of the reference's Op library at thunk level: ARange, Eye, ExtractDiag, Split, Argmax, CumOp.
Each node names the reference Op it replaces; the kernels behind them are `ptk_arange/ptk_argmax/ptk_cumop`
(csrc/ptk_misc.cu) and the strided copy kernels."""

from __future__ import annotations

import numpy as np

from ..runtime import device as dev
from ..runtime import lib as _lib
from .nodes_elemwise import Node
from .values import Val


def _prod(xs) -> int:
    n = 1
    for s in xs:
        n *= int(s)
    return n


class ARangeNode(Node):
    """Reference: ARange, pytensor/tensor/basic.py:3139 (perform = `np.arange(start, stop, step, dtype)`).
    Short ranges are index plumbing and stay host-visible (like Shape_i / MakeVector results); long ones are written by
    `ptk_arange` with NumPy's fill rule `first + i*delta` evaluated in the output type."""

    HOST_MAX = 4096

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype).name
        self.name = f"ARange{{{self.dtype}}}"

    def run(self, vals):
        start, stop, step = (np.asarray(v.host())[()] for v in vals)
        if step == 0:
            np.arange(start, stop, step, dtype=self.dtype)  # raises exactly what the reference's np.arange raises
            raise ValueError("ARange: step must not be zero")
        if all(np.asarray(s).dtype.kind in "iub" for s in (start, stop, step)):
            n = len(range(int(start), int(stop), int(step)))
        else:
            n = max(0, int(np.ceil((float(stop) - float(start)) / float(step))))
        if n <= self.HOST_MAX:
            return [Val(h=np.arange(start, stop, step, dtype=self.dtype))]
        dt = np.dtype(self.dtype)
        first = np.asarray(start).astype(dt)
        second = np.asarray(start + step).astype(dt)
        with np.errstate(over="ignore"):
            delta = (second - first).astype(dt)
        out = dev.empty((n,), self.dtype)
        isf = dt.kind == "f"
        _lib.check(_lib.lib().ptk_arange(_lib.DTYPE_CODE[self.dtype], dev.ptr(out), n,
                                         float(first) if isf else 0.0, float(delta) if isf else 0.0,
                                         0 if isf else int(first.astype(np.int64)), 0 if isf else int(delta.astype(np.int64)),
                                         dev.stream_ptr()), "ptk_arange")
        return [Val(d=out)]


class EyeNode(Node):
    """Reference: Eye, pytensor/tensor/basic.py:1362-1381 (`np.eye(n, m, k, dtype)`): zero fill + ones on diagonal k."""

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype).name
        self.name = f"Eye{{{self.dtype}}}"

    def run(self, vals):
        n, m, k = (int(np.asarray(v.host()).reshape(-1)[0]) for v in vals)
        if n < 0 or m < 0:
            raise ValueError("Eye: negative dimensions are not allowed")
        out = dev.empty((n, m), self.dtype)
        if out.numel():
            _lib.check(_lib.lib().ptk_memset_async(dev.ptr(out), 0, out.numel() * out.element_size(), dev.stream_ptr()),
                       "memset")
            length = min(n, m - k) if k >= 0 else min(n + k, m)
            if length > 0:
                offset = k if k >= 0 else -k * m
                diag = out.as_strided((length,), (m + 1,), offset)
                one = dev.to_device(np.ones((), dtype=self.dtype))
                dev.copy_strided(diag, one.as_strided((length,), (0,)))
        return [Val(d=out)]


class ExtractDiagNode(Node):
    """Reference: ExtractDiag, pytensor/tensor/basic.py:3692-3754 (`x.diagonal(offset, axis1, axis2)`; a view when
    `view=True`, the diagonal axis goes last)."""

    def __init__(self, offset, axis1, axis2, view, name="ExtractDiag"):
        self.offset, self.axis1, self.axis2, self.view = int(offset), int(axis1), int(axis2), bool(view)
        self.views_input0 = self.view  # stream scheduling: the output aliases input 0
        self.name = name

    def run(self, vals):
        v = vals[0]
        if v.d is None:
            return [Val(h=np.array(np.asarray(v.h).diagonal(self.offset, self.axis1, self.axis2)))]
        d = v.d.diagonal(self.offset, self.axis1, self.axis2)  # metadata only
        return [Val(d=d if self.view else dev.clone(d))]


class SplitNode(Node):
    """Reference: Split, pytensor/tensor/basic.py:2237-2283: `np.split` along `axis`; outputs are views of the input
    (view_map :2237), with the reference's three argument checks."""

    views_input0 = True

    def __init__(self, len_splits, axis, name="Split"):
        self.len_splits, self.axis = int(len_splits), int(axis)
        self.name = name

    def run(self, vals):
        x = vals[0]
        splits = np.asarray(vals[1].host()).reshape(-1)
        if len(splits) != self.len_splits:
            raise ValueError("Length of splits is not equal to n_splits")
        if int(splits.sum()) != x.shape[self.axis]:
            raise ValueError(f"Split sizes sum to {int(splits.sum())}; expected {x.shape[self.axis]}")
        if (splits < 0).any():
            raise ValueError("Split sizes cannot be negative")
        outs, pos = [], 0
        if x.d is None:
            h = np.asarray(x.h)
            for s in splits:
                idx = [slice(None)] * h.ndim
                idx[self.axis] = slice(pos, pos + int(s))
                outs.append(Val(h=h[tuple(idx)]))
                pos += int(s)
            return outs
        for s in splits:
            outs.append(Val(d=x.d.narrow(self.axis, pos, int(s))))
            pos += int(s)
        return outs


class ArgmaxNode(Node):
    """Reference: Argmax, pytensor/tensor/math.py:188-206: kept axes first, the reduced axes flattened in their original
    order, `np.argmax` over that last axis -> int64.  When the reduced axes are consecutive the (outer, n, inner) view of
    the contiguous input is used directly; otherwise the reference's transpose is materialised first."""

    def __init__(self, axis, ndim, dtype, name="Argmax"):
        self.axes = tuple(range(ndim)) if axis is None else tuple(sorted(int(a) % ndim for a in axis))
        self.ndim, self.dtype, self.name = ndim, dtype, name

    def run(self, vals):
        t = vals[0].dev()
        shape = tuple(t.shape)
        axes = self.axes
        kept = [i for i in range(self.ndim) if i not in axes]
        oshape = [shape[i] for i in kept]
        if self.ndim == 0:
            return [Val(h=np.asarray(0, dtype="int64"))]
        consecutive = all(b == a + 1 for a, b in zip(axes, axes[1:]))
        if consecutive:
            t = dev.contiguous(t)
            outer, n, inner = _prod(shape[: axes[0]]), _prod(shape[a] for a in axes), _prod(shape[axes[-1] + 1:])
        else:
            t = dev.contiguous(t.permute(kept + list(axes)))
            outer, n, inner = _prod(oshape), _prod(shape[a] for a in axes), 1
        if n == 0 and outer * inner > 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        out = dev.empty(oshape, "int64")
        if out.numel():
            _lib.check(_lib.lib().ptk_argmax(_lib.DTYPE_CODE[self.dtype], dev.ptr(t), dev.ptr(out), outer, n, inner,
                                             dev.stream_ptr()), "ptk_argmax")
        return [Val(d=out)]


class CumOpNode(Node):
    """Reference: CumOp, pytensor/tensor/extra_ops.py:295-321 (`np.cumsum` / `np.cumprod` along `axis`)."""

    SUPPORTED = ("float32", "float64", "int64", "uint64")  # the dtypes np.cumsum / np.cumprod do not widen

    def __init__(self, axis, mode, dtype, name="CumOp"):
        self.axis, self.mode, self.dtype, self.name = int(axis), mode, dtype, name

    def run(self, vals):
        t = dev.contiguous(vals[0].dev())
        shape = tuple(t.shape)
        out = dev.empty(shape, self.dtype)
        if out.numel():
            outer, n, inner = _prod(shape[: self.axis]), shape[self.axis], _prod(shape[self.axis + 1:])
            _lib.check(_lib.lib().ptk_cumop(_lib.DTYPE_CODE[self.dtype], 0 if self.mode == "add" else 1, dev.ptr(t),
                                            dev.ptr(out), outer, n, inner, dev.stream_ptr()), "ptk_cumop")
        return [Val(d=out)]
