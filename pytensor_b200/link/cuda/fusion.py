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


def fuse_steps(steps, output_slots, opts):
    from pytensor_b200.link.cuda.fusion_passes import fuse_elemwise_reduce, fuse_gemm_epilogue, fuse_small_mlp_chains
    from pytensor_b200.link.cuda.fusion_rows import fuse_row_regions

    steps = fuse_row_regions(steps, output_slots, opts)
    steps = fuse_gemm_epilogue(steps, output_slots, opts)
    steps = fuse_small_mlp_chains(steps, output_slots, opts)
    steps = fuse_elemwise_reduce(steps, output_slots, opts)
    return steps
