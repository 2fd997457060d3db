"""General advanced indexing (reference: AdvancedSubtensor pytensor/tensor/subtensor.py:1932, perform :2164 =
`x.__getitem__(indices)`; AdvancedIncSubtensor :2275, perform = `np.add.at` / `x[indices] = y` / `x[indices] += y`).

The fast nodes in nodes_basic.py (TakeNode / PutNode) cover integer index arrays on consecutive axes with every other axis
taken in full.  Everything else NumPy's advanced indexing allows is reduced to that case here, on the device:

  * basic slices mixed into the index are applied first as a strided window (`_probe`, negative steps included) and the
    window is compacted;
  * a boolean mask of m dimensions is turned into the ascending flat positions of its set elements (`ptk_nonzero_*`,
    one device->host read of the COUNT because the output shape depends on it) addressing the m masked axes merged
    into one — exactly `mask.nonzero()` linearised, which is how the reference defines mask indexing (:2046-2051);
  * index arrays separated by a slice ("non-consecutive" advanced indices, :2212-2236) move their axes to the front,
    as NumPy does, by a transposing copy;
  * set / inc through such an index goes through the flat positions of the addressed elements, obtained by pushing an
    element-number array of x's shape through the same gather.

Out-of-range indices raise IndexError at the call's synchronisation point (device error words, nodes_basic.FlagSink)."""

from __future__ import annotations

import numpy as np
import torch

from ..runtime import device as dev
from ..runtime import lib as _lib
from . import nodes_basic as nb
from .nodes_elemwise import Node
from .values import Val


def _nonzero_flat(mask: torch.Tensor) -> torch.Tensor:
    """int64 device vector of the ascending flat positions of the non-zero bytes of `mask` (bool, any shape)."""
    m = dev.contiguous(mask)
    n = m.numel()
    L = _lib.lib()
    wsb = int(L.ptk_nonzero_workspace_bytes(n))
    ws = dev.empty_t((max(wsb // 8, 1),), torch.int64)
    _lib.check(L.ptk_nonzero_count(dev.ptr(m), n, dev.ptr(ws), wsb, dev.stream_ptr()), "ptk_nonzero_count")
    total = int(dev.to_host(ws[-1:])[0])  # raises GraphUnsupported under capture: data-dependent shapes stay eager
    out = dev.empty((total,), "int64")
    if total:
        _lib.check(L.ptk_nonzero_fill(dev.ptr(m), n, dev.ptr(ws), dev.ptr(out), dev.stream_ptr()), "ptk_nonzero_fill")
    return out


class _Plan:
    """What one call's index resolves to: a compact window `xs` of x, the axis group [ax, ax+k) the integer index
    arrays `its` address (already merged for masks, already moved to the front when non-consecutive)."""

    __slots__ = ("xs", "ax", "its", "dims")


def _resolve(x: torch.Tensor, template, kinds, index_vals, name) -> _Plan:
    """template: per indexed axis group either ("s", start, stop, step) with input positions / None, or ("a", pos);
    kinds[pos] = ("bool", ndim) | ("int", ndim)."""
    # 1. the basic part: slices where given, full axes under every advanced index and for the trailing axes
    basic, groups, axis = [], [], 0  # groups: (first axis in the WINDOW, n axes, input position)
    for e in template:
        if e[0] == "s":
            get = lambda p: None if p is None else nb._host_int(index_vals[p])  # noqa: E731
            basic.append(slice(get(e[1]), get(e[2]), get(e[3])))
            axis += 1
        else:
            kind, nd = kinds[e[1]]
            span = nd if kind == "bool" else 1
            groups.append((axis, span, e[1]))
            basic.extend([slice(None)] * span)
            axis += span
    if axis > x.dim():
        raise IndexError("too many indices for array")
    shape, strides, off = nb._probe(tuple(x.shape), tuple(s * x.element_size() for s in x.stride()), basic)
    isz = x.element_size()
    reg = nb._Region(x, shape, tuple(s // isz for s in strides), off // isz)
    t = reg.as_tensor()
    xs = t if (t is not None and t.is_contiguous()) else nb._copy_region_out(reg, None)

    # 2. masks -> flat positions over their merged axes
    new_shape, its, adv_axes, cursor, src_axis = [], [], [], 0, 0
    gi = 0
    while src_axis < xs.dim():
        if gi < len(groups) and groups[gi][0] == src_axis:
            _, span, pos = groups[gi]
            kind, nd = kinds[pos]
            v = index_vals[pos].dev()
            if kind == "bool":
                if tuple(v.shape) != tuple(xs.shape[src_axis:src_axis + span]):
                    raise IndexError(f"boolean index did not match indexed tensor along axis {src_axis}; size of axis is "
                                     f"{tuple(xs.shape[src_axis:src_axis + span])} but size of corresponding boolean axis "
                                     f"is {tuple(v.shape)}")
                its.append(_nonzero_flat(v))
                merged = 1
                for s in xs.shape[src_axis:src_axis + span]:
                    merged *= s
                new_shape.append(merged)
            else:
                its.append(nb._as_int64(v))
                new_shape.append(xs.shape[src_axis])
            adv_axes.append(cursor)
            cursor += 1
            src_axis += span
            gi += 1
        else:
            new_shape.append(xs.shape[src_axis])
            cursor += 1
            src_axis += 1
    xs = xs.reshape(new_shape)

    # 3. non-consecutive advanced axes go to the front (NumPy's rule), as a transposing copy
    if adv_axes != list(range(adv_axes[0], adv_axes[0] + len(adv_axes))):
        perm = adv_axes + [a for a in range(xs.dim()) if a not in adv_axes]
        xs = dev.contiguous(xs.permute(perm))
        adv_axes = list(range(len(adv_axes)))
    p = _Plan()
    p.xs, p.ax, p.its = xs, adv_axes[0], its
    p.dims = [xs.shape[a] for a in adv_axes]
    return p


def _linear(its, dims, name):
    vals = [Val(d=t) for t in its]
    return nb._index_block(vals, dims, name)


def _gather(p: _Plan, name) -> torch.Tensor:
    xs, ax, k = p.xs, p.ax, len(p.its)
    it = _linear(p.its, p.dims, name)
    outer = int(np.prod(xs.shape[:ax], dtype=np.int64))
    inner = int(np.prod(xs.shape[ax + k:], dtype=np.int64))
    n_src = int(np.prod(p.dims, dtype=np.int64))
    out = dev.empty_t(list(xs.shape[:ax]) + list(it.shape) + list(xs.shape[ax + k:]), xs.dtype)
    if out.numel():
        if n_src == 0:
            raise IndexError("index out of bounds (taking from an empty axis)")
        flag = nb._err_flag(f"{name}: index out of bounds")
        _lib.check(_lib.lib().ptk_take(dev.ptr(out), dev.ptr(xs), dev.ptr(it), outer, n_src, it.numel(), inner,
                                       xs.element_size(), flag, dev.stream_ptr()), "ptk_take")
    return out


class AdvIndexNode(Node):
    """x[index] for any mix of slices, integer index arrays (any rank, broadcast against each other) and boolean masks."""

    def __init__(self, template, kinds, name="AdvancedSubtensor"):
        self.template, self.kinds, self.name = template, kinds, name

    def run(self, vals):
        x = vals[0].dev()
        return [Val(d=_gather(_resolve(x, self.template, self.kinds, vals[1:], self.name), self.name))]


class AdvIndexPutNode(Node):
    """x[index] = y / x[index] += y for the same general index.  Duplicate positions: increments accumulate
    (`np.add.at`), unless `ignore_duplicates` (then `x[index] += y` in NumPy's buffered sense: every addressed element
    receives old + y once); sets keep one of the written values (NumPy keeps the last)."""

    def __init__(self, template, kinds, inplace, set_instead_of_inc, ignore_duplicates, dtype, name="AdvancedIncSubtensor"):
        self.template, self.kinds, self.name = template, kinds, name
        self.inplace, self.set_instead_of_inc, self.ignore_duplicates, self.dtype = inplace, set_instead_of_inc, ignore_duplicates, dtype
        self.destroy = {0: 0} if inplace else {}

    def run(self, vals):
        x = vals[0].dev()
        if not self.inplace:
            x = dev.clone(x)
        elif not x.is_contiguous():
            raise NotImplementedError(f"{self.name}: in-place scatter into a non-contiguous buffer")
        x = x if x.is_contiguous() else dev.contiguous(x)
        n = x.numel()
        # element numbers of x, pushed through the gather: the flat position every addressed element lives at
        number = dev.empty(tuple(x.shape), "int64")
        if n:
            _lib.check(_lib.lib().ptk_arange(_lib.DTYPE_CODE["int64"], dev.ptr(number), n, 0.0, 0.0, 0, 1, dev.stream_ptr()),
                       "ptk_arange")
        pos = _gather(_resolve(number, self.template, self.kinds, vals[2:], self.name), self.name)
        y = vals[1].dev()
        if dev.TORCH_TO_NP[y.dtype] != self.dtype:
            from .nodes_cast import cast_to

            y = cast_to(y, self.dtype)
        if y.dim() > pos.dim():
            raise ValueError(f"{self.name}: shape mismatch: value array of shape {tuple(y.shape)} could not be broadcast to "
                             f"indexing result of shape {tuple(pos.shape)}")
        yb = nb._broadcast_view(y, tuple(pos.shape))
        yc = yb if yb.is_contiguous() else dev.contiguous(yb)
        m = pos.numel()
        if not m:
            return [Val(d=x)]
        L = _lib.lib()
        code = _lib.DTYPE_CODE[self.dtype]
        mode = 0 if self.set_instead_of_inc else 1
        if mode == 1 and self.ignore_duplicates:
            # old values + y, then a plain set
            old = dev.empty_t(tuple(pos.shape), x.dtype)
            _lib.check(L.ptk_take(dev.ptr(old), dev.ptr(x), dev.ptr(pos), 1, n, m, 1, x.element_size(),
                                  nb._err_flag(self.name), dev.stream_ptr()), "ptk_take")
            _lib.check(L.ptk_inc_strided(dev.ptr(old), dev.i64_array(old.stride()), dev.ptr(yc), dev.i64_array(yc.stride()),
                                         dev.i64_array(old.shape), old.dim(), code, 1, dev.stream_ptr()), "ptk_inc_strided")
            yc, mode = old, 0
        _lib.check(L.ptk_put(dev.ptr(x), dev.ptr(yc), dev.ptr(pos), 1, n, m, 1, code, mode,
                             nb._err_flag(f"{self.name}: index out of bounds"), dev.stream_ptr()), "ptk_put")
        return [Val(d=x)]
