"""Peephole passes over the lowered step list (see fusion.py)."""

from __future__ import annotations

from pytensor_b200.vm.nodes_elemwise import CAReduceNode, ElemwiseNode, ElemwiseReduceNode
from pytensor_b200.vm.vm import Step


def fuse_elemwise_reduce(steps, output_slots, opts):
    """Elemwise -> CAReduce(trailing axes) of one of its outputs ==> one ElemwiseReduceNode step.

    Safe when: the CAReduce is the ONLY step between the two that touches the Elemwise's outputs or destroys any of
    its inputs (we simply require the CAReduce to be scheduled anywhere later and move it up next to the Elemwise:
    it only reads that one slot), and the reduced slot has no float16/complex dtype issues (same kernels as before).
    """
    readers = {}
    for i, st in enumerate(steps):
        for s in st.ins:
            readers.setdefault(s, []).append(i)
    outset = set(output_slots)
    fused_away = set()
    new_steps = []
    for i, st in enumerate(steps):
        if i in fused_away:
            continue
        impl = st.impl
        if type(impl) is ElemwiseNode:
            cand = None
            for k, s in enumerate(st.outs):
                for j in readers.get(s, []):
                    r = steps[j]
                    if (type(r.impl) is CAReduceNode and j not in fused_away and r.impl.ndim == impl.ndim
                            and impl.ndim > 0 and r.impl.axes
                            and tuple(r.impl.axes) == tuple(range(impl.ndim - len(r.impl.axes), impl.ndim))
                            and r.impl.in_dtype == impl.prog.out_dtypes[k]):
                        cand = (k, s, j)
                        break
                if cand:
                    break
            if cand is not None:
                k, s, j = cand
                others = [x for x in readers.get(s, []) if x != j]
                store = bool(others) or (s in outset)
                node = ElemwiseReduceNode(impl, steps[j].impl, k, store_reduced_input=store)
                new_steps.append(Step(node, st.ins, list(st.outs) + list(steps[j].outs), origin=st.origin))
                fused_away.add(j)
                continue
        new_steps.append(st)
    return new_steps
