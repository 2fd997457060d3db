"""Persistent fused Scan kernel node (K7): an elementwise recurrence runs all its time steps in ONE launch with the
carried state in registers (see codegen/scan.py).  Falls back to the general `ScanNode` whenever the runtime shapes
do not fit the fused kernel's assumptions (all recurrent states and nit-sot outputs share one element shape)."""

from __future__ import annotations

import ctypes
from ctypes import c_int, c_longlong, c_void_p

import numpy as np

from ..codegen import scan as cg_scan
from ..codegen.elemwise import MAX_DIMS
from ..runtime import device as dev
from ..runtime import jit
from ..runtime import lib as _lib
from .nodes_elemwise import Node
from .values import Val


class ScanFusedElemwiseNode(Node):
    def __init__(self, generic, prog, n_seq, state_taps, n_nit, nonseq_leaves, state_ndim, name="ScanFused"):
        """nonseq_leaves: list of (source, index, [view nodes]) with source in {"nonseq", "const"}; `const` carries
        the array itself in `index`."""
        self.generic = generic
        self.prog = prog
        self.n_seq = n_seq
        self.state_taps = [tuple(t) for t in state_taps]
        self.n_nit = n_nit
        self.nonseq_leaves = nonseq_leaves
        self.state_ndim = state_ndim
        self.destroy = dict(generic.destroy)
        self.n_out = generic.n_out
        self.name = name
        self._fn = None
        self._const_cache = {}

    def _kernel(self):
        if self._fn is None:
            self._fn, _ = jit.get_function_gen(
                lambda kn: cg_scan.gen_fused_scan_kernel(self.prog, kn, self.n_seq, self.state_taps, self.n_nit,
                                                         len(self.nonseq_leaves)), "ptk_scan_fused")
        return self._fn

    def run(self, vals):
        g = self.generic
        n_steps = int(np.asarray(vals[0].host()).reshape(-1)[0])
        if n_steps <= 0:
            return g.run(vals)
        n_state = len(self.state_taps)
        seqs = [v.dev() for v in vals[1:g.seqs_arg_offset]]
        for k, s in enumerate(seqs):
            if s.shape[0] < n_steps:
                raise ValueError(f"Sequence {k} has shape {tuple(s.shape)} but the Scan's required number of steps is {n_steps}")
        tap_inputs = [v.dev() for v in vals[g.seqs_arg_offset:g.untraced_arg_offset]]
        nit_lens = [int(np.asarray(v.host()).reshape(-1)[0])
                    for v in vals[g.nit_sot_arg_offset:g.nit_sot_arg_offset + self.n_nit]]
        non_seq_vals = list(vals[g.nit_sot_arg_offset + self.n_nit:])
        # leaves -> device views
        leaves = []
        for src, idx, views in self.nonseq_leaves:
            if src == "nonseq":
                v = non_seq_vals[idx]
            else:
                v = self._const_cache.get(id(idx))
                if v is None:
                    v = Val(h=np.asarray(idx))
                    self._const_cache[id(idx)] = v
            for vn in views:
                v = vn.run([v])[0]
            leaves.append(v.dev())
        D = self.state_ndim
        # element shape S: broadcast of every per-element operand
        shapes = [tuple(t.shape[1:]) for t in tap_inputs] + [tuple(s.shape[1:]) for s in seqs] + [tuple(l.shape) for l in leaves]
        if any(len(s) != D for s in shapes):
            return g.run(vals)
        S = [1] * D
        for s in shapes:
            for k in range(D):
                if s[k] != 1:
                    if S[k] != 1 and S[k] != s[k]:
                        return g.run(vals)
                    S[k] = s[k]
        if any(tuple(t.shape[1:]) != tuple(S) for t in tap_inputs):
            return g.run(vals)
        total = 1
        for s in S:
            total *= s
        store = [int(t.shape[0]) for t in tap_inputs] + nit_lens
        L = [-min(t) for t in self.state_taps]
        if any(st < l for st, l in zip(store, L)) or any(n <= 0 for n in nit_lens):
            return g.run(vals)
        outs = [None] * g.n_out
        for idx in range(n_state):
            outs[idx] = tap_inputs[idx] if idx in self.destroy else dev.clone(tap_inputs[idx])
        for k in range(self.n_nit):
            outs[n_state + k] = dev.empty((nit_lens[k], *S), g.out_dtypes[n_state + k])
        if total == 0:
            return [Val(d=o) for o in outs]
        if D > MAX_DIMS or n_steps >= 2 ** 31 or max(store) >= 2 ** 31:  # the kernel counts steps / slots in int32
            return g.run(vals)
        ops = seqs + outs[:n_state] + outs[n_state:n_state + self.n_nit] + leaves
        nops = len(ops)

        class ScDims(ctypes.Structure):
            _fields_ = [("ndim", c_int), ("shape", c_longlong * MAX_DIMS), ("st", (c_longlong * MAX_DIMS) * max(nops, 1)),
                        ("tstride", c_longlong * max(nops, 1)), ("store", c_longlong * max(n_state + self.n_nit, 1))]

        d = ScDims()
        d.ndim = D
        for k in range(D):
            d.shape[k] = S[k]
        for j, t in enumerate(ops):
            per_elem = j >= self.n_seq + n_state + self.n_nit  # leaves have no time dim
            for k in range(D):
                dim = k if per_elem else k + 1
                d.st[j][k] = 0 if t.shape[dim] == 1 else t.stride(dim)
            d.tstride[j] = 0 if per_elem else t.stride(0)
        for k, s in enumerate(store):
            d.store[k] = s
        args = [c_void_p(dev.ptr(t)) for t in ops] + [d, c_longlong(total), c_longlong(n_steps)]
        grid = min(max(1, (total + 255) // 256), _lib.sm_count() * 32)
        jit.launch(self._kernel(), (grid,), (256,), jit.KernelArgs(args), 0, dev.stream_ptr())
        mintaps = [min(t) for t in self.state_taps] + [0] * self.n_nit
        pos = [((-mintaps[k]) + n_steps) % store[k] for k in range(n_state + self.n_nit)]
        g.finalize(outs, store, pos, n_steps, n_steps, first=0, last=n_state + self.n_nit, mintaps=mintaps)
        return [Val(d=o) for o in outs]
