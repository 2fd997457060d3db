"""Scan on the device (K7, general path): the time loop of `Scan` with every buffer resident in HBM.

Semantics restated from the reference loop (`Scan.perform`, pytensor/scan/op.py:1827-2329, and its Cython twin
pytensor/scan/scan_perform.pyx:76-603):
  * outer inputs  [n_steps, seqs.., mit_mot.., mit_sot.., sit_sot.., untraced.., nit_sot lengths.., non_seqs..]
    (op.py:1830-1847); outer outputs [mit_mot.., mit_sot.., sit_sot.., nit_sot.., untraced..];
  * tap buffers are circular: pos = (-mintap) mod store_steps (op.py:1931-1934), step i reads taps at
    (pos + t) mod store_steps (:1968-1990) and writes the new value at pos (:2152-2190), mit-mot writes at
    pos + out_slice (:2133-2141), then pos advances (:2246-2248);
  * after the loop the circular buffers are rotated into chronological order (:2253-2282); `until` truncates
    (:2289-2305); n_steps == 0 returns empty nit-sots and untraced inputs unchanged (:1907-1921).
The inner graph is itself a lowered CUDA Program executed once per step — no host math, no host copies of tensor data;
`until` conditions are the only per-step device->host read.  The fused persistent kernel for elementwise recurrences is
`ScanFusedElemwiseNode` (nodes_scan_fused.py); this node is the general fallback.
"""

from __future__ import annotations

import numpy as np

from ..runtime import device as dev
from .nodes_elemwise import Node
from .values import Val


class ScanNode(Node):
    def __init__(self, info: dict, inner_program, out_dtypes, out_ndims, destroy, name="Scan"):
        """info: plain-data copy of ScanInfo (n_seqs, mit_mot_in_slices, mit_mot_out_slices, mit_sot_in_slices,
        sit_sot_in_slices, n_nit_sot, n_untraced_sit_sot, n_non_seqs, as_while)."""
        from .vm import Executor

        self.info = info
        self.inner = Executor(inner_program, allow_gc=True)
        self.out_dtypes = list(out_dtypes)
        self.out_ndims = list(out_ndims)
        self.destroy = dict(destroy)
        self.name = name
        i = info
        self.n_mit_mot = len(i["mit_mot_in_slices"])
        self.n_mit_sot = len(i["mit_sot_in_slices"])
        self.n_sit_sot = len(i["sit_sot_in_slices"])
        self.n_tap_outs = self.n_mit_mot + self.n_mit_sot + self.n_sit_sot
        self.tap_array = list(i["mit_mot_in_slices"]) + list(i["mit_sot_in_slices"]) + list(i["sit_sot_in_slices"])
        self.mintaps = [min(t) for t in self.tap_array] + [0] * i["n_nit_sot"]
        self.seqs_arg_offset = 1 + i["n_seqs"]
        self.untraced_arg_offset = self.seqs_arg_offset + self.n_tap_outs
        self.nit_sot_arg_offset = self.untraced_arg_offset + i["n_untraced_sit_sot"]
        self.n_out = self.n_tap_outs + i["n_nit_sot"] + i["n_untraced_sit_sot"]

    def run(self, vals):
        info = self.info
        n_steps = int(np.asarray(vals[0].host()).reshape(-1)[0])
        if n_steps < 0:
            raise IndexError(f"Scan was asked to run for negative number of step {n_steps}")
        seqs = [v.dev() for v in vals[1:self.seqs_arg_offset]]
        for k, s in enumerate(seqs):
            if s.shape[0] < n_steps:
                raise ValueError(f"Sequence {k} has shape {tuple(s.shape)} but the Scan's required number of steps is {n_steps}")
        n_nit = info["n_nit_sot"]
        n_unt = info["n_untraced_sit_sot"]
        tap_inputs = vals[self.seqs_arg_offset:self.untraced_arg_offset]
        store_steps = [int(v.shape[0]) for v in tap_inputs]
        store_steps += [int(np.asarray(v.host()).reshape(-1)[0])
                        for v in vals[self.nit_sot_arg_offset:self.nit_sot_arg_offset + n_nit]]
        # output buffers for the tap outputs: in place when the op owns the input buffer, a copy otherwise
        outs = [None] * self.n_out
        for idx in range(self.n_tap_outs):
            t = tap_inputs[idx].dev()
            outs[idx] = t if idx in self.destroy else dev.clone(t)
        untraced = [v for v in vals[self.untraced_arg_offset:self.untraced_arg_offset + n_unt]]
        non_seqs = list(vals[self.nit_sot_arg_offset + n_nit:])
        nit_end = self.n_tap_outs + n_nit
        if n_steps == 0:
            for j in range(self.n_tap_outs, nit_end):
                outs[j] = dev.empty((0,) * self.out_ndims[j], self.out_dtypes[j])
            for j in range(n_unt):
                outs[nit_end + j] = untraced[j].dev()
            return [Val(d=o) for o in outs]

        pos = [(-self.mintaps[idx]) % store_steps[idx] for idx in range(self.n_tap_outs + n_nit)]
        n_mm_outs = sum(len(s) for s in info["mit_mot_out_slices"])
        i = 0
        cond = True
        cur_untraced = untraced
        while i < n_steps and cond:
            inner_in = [Val(d=s[i]) for s in seqs]
            for idx, taps in enumerate(self.tap_array):
                for t in taps:
                    inner_in.append(Val(d=outs[idx][(pos[idx] + t) % store_steps[idx]]))
            inner_in.extend(cur_untraced)
            inner_in.extend(non_seqs)
            res = self.inner.run(inner_in)
            if info["as_while"]:
                cond = bool(np.asarray(res[-1].host()).reshape(-1)[0] == 0)
            # mit-mot outputs: written at pos + out_slice (no wrap-around; op.py:2133-2141)
            k = 0
            for g, out_slices in enumerate(info["mit_mot_out_slices"]):
                for sl in out_slices:
                    dev.copy_strided(outs[g][sl + pos[g]], res[k].dev())
                    k += 1
            # mit-sot / sit-sot
            for j in range(self.n_mit_mot, self.n_tap_outs):
                r = res[n_mm_outs + (j - self.n_mit_mot)].dev()
                dst = outs[j][pos[j]]
                if tuple(dst.shape) != tuple(r.shape):
                    raise ValueError("An output of the Scan has changed shape.")
                dev.copy_strided(dst, r)
            # nit-sot
            for j in range(self.n_tap_outs, nit_end):
                r = res[n_mm_outs + (j - self.n_mit_mot)].dev()
                if i == 0:
                    outs[j] = dev.empty((store_steps[j], *r.shape), self.out_dtypes[j])
                dev.copy_strided(outs[j][pos[j]], r)
            # untraced sit-sot: carried as values, never stored
            base = n_mm_outs + (nit_end - self.n_mit_mot)
            cur_untraced = [res[base + j] for j in range(n_unt)]
            pos = [(p + 1) % s for p, s in zip(pos, store_steps)]
            i += 1
        for j in range(n_unt):
            outs[nit_end + j] = cur_untraced[j].dev()

        self.finalize(outs, store_steps, pos, i, n_steps, first=self.n_mit_mot, last=nit_end, mintaps=self.mintaps)
        return [Val(d=o) for o in outs]

    def finalize(self, outs, store_steps, pos, i, n_steps, first, last, mintaps):
        """Chronological re-ordering of the circular buffers / zero-fill / `until` truncation (op.py:2250-2305)."""
        for idx in range(first, last):
            st = store_steps[idx]
            buf = outs[idx]
            if st < i - mintaps[idx] and pos[idx] < st:
                pdx = pos[idx]
                if pdx != 0:
                    new = dev.empty_t(buf.shape, buf.dtype)
                    dev.copy_strided(new[: st - pdx], buf[pdx:])
                    dev.copy_strided(new[st - pdx:], buf[:pdx])
                    outs[idx] = new
            elif st > i - mintaps[idx]:
                tail = buf[i - mintaps[idx]:]
                if tail.numel():
                    from ..runtime import lib as _lib

                    if tail.is_contiguous():
                        _lib.check(_lib.lib().ptk_memset_async(dev.ptr(tail), 0, tail.numel() * tail.element_size(),
                                                               dev.stream_ptr()), "memset")
                    else:
                        zero = dev.to_device(np.zeros((), dtype=self.out_dtypes[idx]))
                        dev.copy_strided(tail, zero.as_strided(tuple(tail.shape), (0,) * tail.dim()))
                if i < n_steps:
                    outs[idx] = buf[: buf.shape[0] - (n_steps - i)]
