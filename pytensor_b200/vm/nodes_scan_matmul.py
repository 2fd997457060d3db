"""Scan whose inner graph is ONE matrix product of the carried state: h <- act(h @ W + b) (BASELINE.json configs[3],
secondary recurrence).  Reference loop being replaced: pytensor/scan/scan_perform.pyx:311-541 — per step
it juggles the tap cells, calls the inner VM (Dot22 -> sgemm_, then the Composite) and copies the result into the
circular buffer.

Here the loop is a chain of tensor-core launches and nothing else:
  * W (a non-sequence: it cannot change during the loop) is staged into the kernel's operand layout ONCE;
  * the state is staged once, before the first step; from then on the epilogue of step t writes, next to the fp32 result,
    the staged operand of step t+1 (bf16, or the three-piece split of the fp32-accurate mode);
  * every step's result is written by the kernel's epilogue straight into its slot of the circular tap buffer — no
    per-step copy, no inner executor, no host work besides the launch; under the VM's CUDA-graph capture the whole loop
    becomes kernel nodes of the outer graph.
Small / fp64 recurrences take the FMA kernel per step, also writing in place.  Buffer protocol (circular positions, final
re-ordering, zero fill) is the general ScanNode's (nodes_scan.py), which also serves n_steps == 0 and one-slot buffers."""

from __future__ import annotations

import numpy as np

from ..runtime import device as dev
from . import nodes_blas as nblas
from .nodes_elemwise import Node
from .values import Val


class ScanMatmulRecurrenceNode(Node):
    def __init__(self, generic, w_src, bias_src, act, precision, dtype, name="Scan"):
        """w_src / bias_src: ("in", index into the node's inputs) | ("const", ndarray) | None."""
        self.generic = generic
        self.w_src, self.bias_src, self.act, self.precision, self.dtype = w_src, bias_src, act, precision, dtype
        self.destroy = dict(generic.destroy)
        self.name = name

    def _operand(self, src, vals):
        if src is None:
            return None
        return vals[src[1]].dev() if src[0] == "in" else dev.to_device(np.asarray(src[1]))

    def run(self, vals):
        g = self.generic
        n_steps = int(np.asarray(vals[0].host()).reshape(-1)[0])
        buf = vals[g.seqs_arg_offset].dev()
        S = int(buf.shape[0])
        W = self._operand(self.w_src, vals)
        if n_steps <= 0 or S < 2 or buf.dim() != 3 or W.dim() != 2:
            return g.run(vals)
        M, N = int(buf.shape[1]), int(buf.shape[2])
        if tuple(W.shape) != (N, N):
            raise ValueError(f"{self.name}: the state {(M, N)} times W {tuple(W.shape)} changes shape")
        bias = self._operand(self.bias_src, vals)
        if bias is not None:
            if bias.shape[-1] != N or bias.numel() != N:
                return g.run(vals)
            bias = bias.reshape(-1) if bias.is_contiguous() else dev.contiguous(bias).reshape(-1)
        out = buf if 0 in self.destroy else dev.clone(buf)
        if not out.is_contiguous():
            return g.run(vals)
        pos = 1 % S  # (-mintap) mod store_steps with the single tap -1 (scan/op.py:1931-1934)
        plan = nblas.tc_plan(self.dtype, self.precision, M, N, N) if M * N else None
        if plan is not None:
            pieces, terms = plan
            Wst = nblas.stage_operand(W, pieces, transposed=True)
            cur = nblas.stage_operand(out[(pos - 1) % S], pieces)
            # the epilogue writes the next step's operand when it can: always in the bf16 mode; in the fp32-accurate mode
            # with error-free leading pieces only for a bounded activation (tanh) — otherwise one staging pass per step
            chain = pieces == 1 or nblas.can_chain_pieces(self.act)
            nxt = nblas.Staged(M, N, pieces, aligned=cur.aligned) if chain else None
            for _ in range(n_steps):
                nblas.gemm_staged(cur, Wst, terms, 1.0, 0.0, out[pos], bias=bias, act=self.act, out=nxt)
                if chain:
                    cur, nxt = nxt, cur
                else:
                    cur = nblas.stage_operand(out[pos], pieces)
                pos = (pos + 1) % S
        else:
            for _ in range(n_steps):
                if M * N:
                    nblas.gemm(self.dtype, 1.0, out[(pos - 1) % S], W, 0.0, out[pos], 0, bias=bias, act=self.act)
                pos = (pos + 1) % S
        outs = [out]
        g.finalize(outs, [S], [pos], n_steps, n_steps, first=0, last=1, mintaps=g.mintaps)
        return [Val(d=outs[0])]
