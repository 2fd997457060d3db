"""Lowering of `Scan` (pytensor/scan/op.py:839): the inner graph is lowered with the same per-node dispatch into an
inner CUDA Program; the time loop runs in `ScanNode` (general) or in one persistent fused kernel
(`ScanFusedElemwiseNode`) when the inner graph is a pure elementwise recurrence.  `Scan.make_thunk` (op.py:1621) and
the Cython loop (scan_perform.pyx) are never used."""

from __future__ import annotations

from pytensor_b200.vm.nodes_scan import ScanNode


def scan_info_dict(info):
    return {
        "n_seqs": int(info.n_seqs),
        "mit_mot_in_slices": [tuple(int(t) for t in s) for s in info.mit_mot_in_slices],
        "mit_mot_out_slices": [tuple(int(t) for t in s) for s in info.mit_mot_out_slices],
        "mit_sot_in_slices": [tuple(int(t) for t in s) for s in info.mit_sot_in_slices],
        "sit_sot_in_slices": [tuple(int(t) for t in s) for s in info.sit_sot_in_slices],
        "n_nit_sot": int(info.n_nit_sot),
        "n_untraced_sit_sot": int(info.n_untraced_sit_sot),
        "n_non_seqs": int(info.n_non_seqs),
        "as_while": bool(info.as_while),
    }


def lower_scan(node, opts):
    from pytensor_b200.link.cuda.linker import build_program

    op = node.op
    inner = op.fgraph
    order = list(inner.toposort())
    inner_opts = dict(opts)
    program, _ = build_program(inner, order, inner_opts, storage_map=None)
    info = scan_info_dict(op.info)
    destroy = {int(o): int(i[0]) for o, i in getattr(op, "destroy_map", {}).items()}
    out_dtypes = [o.type.dtype for o in node.outputs]
    out_ndims = [o.type.ndim for o in node.outputs]
    generic = ScanNode(info, program, out_dtypes, out_ndims, destroy, name=str(op))
    if opts.get("fuse", True):
        rec = try_matmul_recurrence(node, info, program, generic, opts)
        if rec is not None:
            return rec
        try:
            from pytensor_b200.link.cuda.lower_scan_fused import try_fused_elemwise_scan

            fused = try_fused_elemwise_scan(node, info, program, generic)
            if fused is not None:
                return fused
        except ImportError:
            pass
    return generic


def try_matmul_recurrence(node, info, program, generic, opts):
    """h <- act(h @ W + b) with one sit-sot state and nothing else in the loop: ScanMatmulRecurrenceNode, else None."""
    from pytensor_b200.vm.nodes_scan_matmul import ScanMatmulRecurrenceNode

    if (info["n_seqs"] or info["mit_mot_in_slices"] or info["mit_sot_in_slices"] or info["sit_sot_in_slices"] != [(-1,)]
            or info["n_nit_sot"] or info["n_untraced_sit_sot"] or info["as_while"]):
        return None
    if len(program.steps) != 1 or len(program.inputs) != 1 + info["n_non_seqs"]:
        return None
    st = program.steps[0]
    kind = type(st.impl).__name__
    if kind not in ("GemmBiasActNode", "Dot22Node") or getattr(st.impl, "scalar", False):
        return None
    if list(program.outputs) != [st.outs[0]] or st.ins[0] != program.inputs[0] or program.inputs[0] in st.ins[1:]:
        return None
    dtype = st.impl.dtype
    if dtype not in ("float32", "float64") or node.outputs[0].type.ndim != 3:
        return None
    first_non_seq = generic.nit_sot_arg_offset + info["n_nit_sot"]

    def source(slot):
        if slot in program.constants:
            return ("const", program.constants[slot])
        if slot in program.inputs:
            return ("in", first_non_seq + program.inputs.index(slot) - 1)
        return False

    w_src = source(st.ins[1])
    bias_src = source(st.ins[2]) if (kind == "GemmBiasActNode" and len(st.ins) > 2) else None
    if w_src is False or bias_src is False:
        return None
    act = st.impl.act if kind == "GemmBiasActNode" else 0
    return ScanMatmulRecurrenceNode(generic, w_src, bias_src, act, st.impl.precision, dtype, name=str(node.op))
