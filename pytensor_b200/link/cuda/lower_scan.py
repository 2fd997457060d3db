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
        try:
            from pytensor_b200.link.cuda.lower_scan_fused import try_fused_elemwise_scan

            fused = try_fused_elemwise_scan(node, info, program, generic)
            if fused is not None:
                return fused
        except ImportError:
            pass
    return generic
