"""Linker-level kernel-boundary re-fusion (peepholes over the lowered step list).

The reference's rewriter decides the node boundaries we receive; two fusions it cannot express are done here, after
lowering, because they only make sense for a device backend:
  * Elemwise(Composite) -> CAReduce over the trailing axes of one of its outputs, as ONE kernel (K3): the reference
    refuses to fuse a multi-input Elemwise into a CAReduce (pytensor/tensor/rewriting/elemwise.py:1119-1121), so the
    C linker re-reads the elementwise result from memory; the fused kernel keeps it in registers.
  * Dot22 -> Elemwise{act(x + bias)} epilogue (K5), see nodes_blas; runs of such layers as one launch when they are small
    (MlpChainNode).
  * whole sub-graphs over independent batch rows (gather -> skinny Gemm -> Elemwise -> Sum / Gemm / scatter-add -> sums
    over the batch) as ONE persistent kernel: fusion_rows.py / codegen/rowfuse.py.
"""

from __future__ import annotations


def read_before_write(steps, defined) -> int | None:
    """Index of the first step that reads a slot nobody has written yet (`defined`: the slots that hold a value before the
    first step — inputs, constants, caller-filled cells), or None for a well-formed list."""
    have = set(defined)
    for k, st in enumerate(steps):
        if any(s not in have for s in st.ins):
            return k
        have.update(st.outs)
    return None


def fuse_steps(steps, output_slots, opts):
    from pytensor_b200.link.cuda.fusion_passes import fuse_elemwise_reduce, fuse_gemm_epilogue, fuse_small_mlp_chains
    from pytensor_b200.link.cuda.fusion_rows import fuse_row_regions

    # Every pass re-orders or replaces steps; each result is checked for the one property execution depends on — a step
    # only reads what an earlier step (or the caller) has written, and every function output is written — and a pass whose
    # result violates it is dropped (the program stays correct, just less fused) instead of failing at run time.
    produced = {s for st in steps for s in st.outs}
    defined = {s for st in steps for s in st.ins if s not in produced} | set(opts.get("constants", {}))
    for fuse_pass in (fuse_row_regions, fuse_gemm_epilogue, fuse_small_mlp_chains, fuse_elemwise_reduce):
        new = fuse_pass(steps, output_slots, opts)
        written = defined | {s for st in new for s in st.outs}
        if read_before_write(new, defined) is None and all(s in written for s in output_slots):
            steps = new
        else:
            import warnings

            warnings.warn(f"pytensor_b200: {fuse_pass.__name__} produced an ill-ordered step list; pass skipped", RuntimeWarning)
    return steps
