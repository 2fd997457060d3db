"""dtype conversion on the device (an Elemwise{Cast} kernel; pytensor/scalar/basic.py:2435-2503)."""

from __future__ import annotations

from ..codegen.scalar import single_op_program
from ..runtime import device as dev
from .values import Val

_cache = {}


def cast_to(t, dtype: str):
    src = dev.TORCH_TO_NP[t.dtype]
    if src == dtype:
        return t
    from .nodes_elemwise import ElemwiseNode

    key = (src, dtype, t.dim())
    node = _cache.get(key)
    if node is None:
        node = ElemwiseNode(single_op_program("Cast", [src], dtype), t.dim(), [(False,) * t.dim()], {}, name=f"Cast{{{dtype}}}")
        _cache[key] = node
    return node.run([Val(d=t)])[0].d
