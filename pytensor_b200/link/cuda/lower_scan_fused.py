"""Recognise Scans whose inner graph is a pure elementwise recurrence and lower them to the persistent fused kernel.

The analysis runs on the already-lowered inner Program (Elemwise / DimShuffle steps only), composing the per-node
ScalarPrograms into one program: inputs [seq slices, state taps, broadcast non-seq leaves] -> outputs [new states,
nit-sot values] — "the inner graph inlined" of BASELINE.json's north_star."""

from __future__ import annotations

import numpy as np

from pytensor_b200.codegen.scalar import ScalarInst, ScalarProgram, emit_body
from pytensor_b200.vm.nodes_basic import DimShuffleNode, ViewNode
from pytensor_b200.vm.nodes_elemwise import ElemwiseNode
from pytensor_b200.vm.nodes_scan_fused import ScanFusedElemwiseNode


def try_fused_elemwise_scan(node, info, program, generic):
    if info["mit_mot_in_slices"] or info["as_while"] or info["n_untraced_sit_sot"]:
        return None
    n_seq = info["n_seqs"]
    state_taps = list(info["mit_sot_in_slices"]) + list(info["sit_sot_in_slices"])
    n_state = len(state_taps)
    n_nit = info["n_nit_sot"]
    n_taps = sum(len(t) for t in state_taps)
    inner_inputs = node.op.fgraph.inputs
    D = None
    if n_state:
        D = inner_inputs[n_seq].type.ndim
    elif n_seq:
        D = inner_inputs[0].type.ndim
    else:
        return None
    in_dtypes = [v.type.dtype for v in inner_inputs[: n_seq + n_taps]]
    out_dtypes = [v.type.dtype for v in node.op.fgraph.outputs]
    if any(v.type.ndim != D for v in inner_inputs[: n_seq + n_taps]) or any(
        v.type.ndim != D for v in node.op.fgraph.outputs
    ):
        return None
    prog = ScalarProgram(in_dtypes=list(in_dtypes), out_dtypes=list(out_dtypes))
    nonseq_leaves = []  # (source, index/array, [view nodes])
    leaf_input = {}     # leaf id -> program input index
    sym = {}            # slot -> ("ref", ref, dtype) | ("leaf", leaf_id, dtype, ndim)

    def add_leaf(source, index, views, dtype):
        nonseq_leaves.append((source, index, list(views)))
        return len(nonseq_leaves) - 1

    for k, s in enumerate(program.inputs):
        v = inner_inputs[k]
        if k < n_seq + n_taps:
            sym[s] = ("ref", ("i", k), v.type.dtype)
        else:
            lid = add_leaf("nonseq", k - (n_seq + n_taps), [], v.type.dtype)
            sym[s] = ("leaf", lid, v.type.dtype, v.type.ndim)
    for s, arr in program.constants.items():
        arr = np.asarray(arr)
        if arr.size == 1:
            prog.consts.append((arr.dtype.name, arr.reshape(-1)[0].item()))
            sym[s] = ("ref", ("c", len(prog.consts) - 1), arr.dtype.name)
        else:
            lid = add_leaf("const", arr, [], arr.dtype.name)
            sym[s] = ("leaf", lid, arr.dtype.name, arr.ndim)

    def as_ref(entry):
        if entry[0] == "ref":
            return entry[1], entry[2]
        _, lid, dtype, ndim = entry
        if ndim != D:
            raise _NoFuse()
        if lid not in leaf_input:
            leaf_input[lid] = len(prog.in_dtypes)
            prog.in_dtypes.append(dtype)
        return ("i", leaf_input[lid]), dtype

    class _NoFuse(Exception):
        pass

    try:
        for st in program.steps:
            impl = st.impl
            if isinstance(impl, DimShuffleNode | ViewNode):
                src = sym.get(st.ins[0])
                if src is None or src[0] != "leaf":
                    if isinstance(impl, ViewNode) and src is not None:
                        sym[st.outs[0]] = src
                        continue
                    return None
                _, lid, dtype, ndim = src
                source, index, views = nonseq_leaves[lid]
                new_lid = add_leaf(source, index, [*views, impl], dtype)
                new_ndim = len(impl.new_order) if isinstance(impl, DimShuffleNode) else ndim
                sym[st.outs[0]] = ("leaf", new_lid, dtype, new_ndim)
            elif isinstance(impl, ElemwiseNode):
                if impl.ndim != D:
                    return None
                arg = [as_ref(sym[s]) for s in st.ins]
                base_c, base_t = len(prog.consts), len(prog.insts)
                prog.consts.extend(impl.prog.consts)

                def remap(r):
                    kind, k = r
                    if kind == "i":
                        return arg[k][0]
                    if kind == "c":
                        return ("c", base_c + k)
                    return ("t", base_t + k)

                for inst in impl.prog.insts:
                    prog.insts.append(ScalarInst(inst.op, [remap(r) for r in inst.args], list(inst.in_dtypes),
                                                 inst.out_dtype))
                for s, r, dt in zip(st.outs, impl.prog.outputs, impl.prog.out_dtypes):
                    sym[s] = ("ref", remap(r), dt)
            else:
                return None
        outs = []
        for s, dt in zip(program.outputs, out_dtypes):
            ref, rdt = as_ref(sym[s])
            if ref[0] != "t" or rdt != dt:
                prog.insts.append(ScalarInst("Cast" if rdt != dt else "Identity", [ref], [rdt], dt))
                ref = ("t", len(prog.insts) - 1)
            outs.append(ref)
        prog.outputs = outs
    except (_NoFuse, KeyError):
        return None
    # keep only the leaves the program really reads, in program-input order
    used = sorted(leaf_input.items(), key=lambda kv: kv[1])
    leaves = [nonseq_leaves[lid] for lid, _ in used]
    if len(prog.in_dtypes) != n_seq + n_taps + len(leaves):
        return None
    try:
        emit_body(prog)
    except NotImplementedError:
        return None
    return ScanFusedElemwiseNode(generic, prog, n_seq, state_taps, n_nit, leaves, D, name=f"{node.op}[fused-persistent]")
