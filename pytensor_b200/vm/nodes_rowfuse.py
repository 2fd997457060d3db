"""Row-fused region node: a whole sub-graph over independent batch rows as ONE persistent kernel + one tiny finishing
kernel (codegen/rowfuse.py).  Replaces, for PyMC-style batched logp+grad graphs, the chain of reference thunks
AdvancedSubtensor (pytensor/tensor/subtensor.py:1932) -> Gemm (blas/gemm.py:76) -> Elemwise (elemwise.py:375) -> Sum
(elemwise.py:1233) / AdvancedIncSubtensor (subtensor.py:2275) that the C linker runs one after the other over
materialised (B, n) arrays.

The node always keeps the region's constituent steps: whenever the runtime shapes / layouts do not fit the fused kernel
(too few rows, no skinny side on a matrix product, shared-memory budget, dtype surprises) it runs them one by one on the
device — same results, same errors as the unfused program.  No pytensor import (programs pickle)."""

from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_longlong, c_void_p

import numpy as np

from ..codegen import rowfuse as cg
from ..runtime import device as dev
from ..runtime import jit
from ..runtime import lib as _lib
from . import nodes_basic
from .nodes_elemwise import Node
from .values import Val

MIN_ROWS = int(os.environ.get("PTK_ROWFUSE_MIN_ROWS", "64"))


class RowRegionNode(Node):
    def __init__(self, plan: cg.RegionPlan, sub_steps, ext_slots, out_slots, destroy=None, name="RowRegion", consts=None):
        """`sub_steps`: vm.Step list over program slot ids (the unfused path); `ext_slots` / `out_slots`: the slot ids
        this node's inputs / outputs stand for, in order; `consts`: slot -> array of the scalar constants that the fused
        kernel has baked into its source (the unfused path still reads them as values)."""
        self.plan = plan
        self.consts = {int(k): np.asarray(v) for k, v in (consts or {}).items()}
        self.sub_steps = list(sub_steps)
        self.ext_slots = list(ext_slots)
        self.out_slots = list(out_slots)
        self.n_out = len(out_slots)
        self.destroy = dict(destroy or {})
        self.name = name
        self._kernels = {}
        self.fused_calls = 0
        self.unfused_calls = 0
        self.last_reason = None
        last = {}
        for i, st in enumerate(self.sub_steps):
            for s in st.ins:
                last[s] = i
        keep = set(self.out_slots) | set(self.ext_slots) | set(self.consts)
        self._free = [[s for s, j in last.items() if j == i and s not in keep] for i in range(len(self.sub_steps))]

    def __repr__(self):
        return f"<RowRegionNode {self.name}: {len(self.plan.ops)} ops fused[rowfuse]>"

    # ---- unfused path ------------------------------------------------------------------------------------------------
    def _unfused(self, vals):
        self.unfused_calls += 1
        slots = {k: Val(h=v) for k, v in self.consts.items()}
        slots.update(zip(self.ext_slots, vals))
        for i, st in enumerate(self.sub_steps):
            res = st.impl.run([slots[j] for j in st.ins])
            for j, r in zip(st.outs, res):
                slots[j] = r
            for j in self._free[i]:
                slots.pop(j, None)
        return [slots[s] for s in self.out_slots]

    # ---- fused path ----------------------------------------------------------------------------------------------------
    def _bind(self, vals):
        """Check the runtime operands against the plan; returns (B, dims, tensors) or raises NotFusable."""
        plan = self.plan
        dims = {}
        B = None

        def unify(sym, n):
            if dims.setdefault(sym, int(n)) != int(n):
                raise cg.NotFusable("domain sizes disagree")

        tens = [None] * plan.n_ext
        for v in plan.vals:
            if v.ext < 0 or v.const is not None:
                continue
            val = vals[v.ext]
            if val.dtype != v.dtype:
                raise cg.NotFusable(f"dtype {val.dtype} != {v.dtype}")
            shp = val.shape
            if v.kind == "R1":
                if len(shp) != 2:
                    raise cg.NotFusable("rank")
                b = shp[0]
                unify(v.dom, shp[1])
            elif v.kind == "R0":
                if len(shp) not in (1, 2) or (len(shp) == 2 and shp[1] != 1):
                    raise cg.NotFusable("rank")
                b = shp[0]
            elif v.kind == "S1":
                if len(shp) == 2 and shp[0] == 1:
                    unify(v.dom, shp[1])
                elif len(shp) == 1:
                    unify(v.dom, shp[0])
                else:
                    raise cg.NotFusable("rank")
                b = None
            elif v.kind == "S2":
                if len(shp) != 2:
                    raise cg.NotFusable("rank")
                unify(v.dom, shp[0])
                unify(v.dom2, shp[1])
                b = None
            else:  # S0
                n = 1
                for s in shp:
                    n *= s
                if n != 1:
                    raise cg.NotFusable("scalar operand with more than one element")
                b = None
            if b is not None:
                if B is None:
                    B = int(b)
                elif B != int(b):
                    raise cg.NotFusable("batch sizes disagree")
            tens[v.ext] = val.dev()
        if B is None or B < MIN_ROWS:
            raise cg.NotFusable(f"batch of {B} rows")
        for v in plan.vals:
            if v.kind in ("R1", "S1") and v.dom not in dims:
                raise cg.NotFusable("a domain size is not determined by the inputs")
            if v.kind in ("R1", "S1") and not 1 <= dims[v.dom] <= 65536:
                raise cg.NotFusable("domain size")
        return B, dims, tens

    def _layouts(self, tens, ptr_of=None):
        """ExtLayout per input + the list of distinct base pointers (inputs that are views of the same memory — X and
        X.T — share one kernel parameter).  `ptr_of`: address of a tensor-like (tests hand in NumPy shims)."""
        groups = {}
        lay = [None] * self.plan.n_ext
        if ptr_of is None:
            ptr_of = (lambda t: 4096 * (1 + id(t) % 1000003)) if _lib.TRACE_ONLY else dev.ptr
        for v in self.plan.vals:
            if v.ext < 0 or v.const is not None or lay[v.ext] is not None:
                continue
            t = tens[v.ext]
            p = ptr_of(t)
            key = (p, v.dtype)
            g = groups.setdefault(key, len(groups))
            if v.kind == "R1":
                st = (t.stride(0), t.stride(1))
            elif v.kind == "R0":
                st = (t.stride(0),)
            elif v.kind == "S1":
                st = (t.stride(-1),)
            elif v.kind == "S2":
                st = (t.stride(0), t.stride(1))
            else:
                st = (0,)
            lay[v.ext] = cg.ExtLayout(g, tuple(int(s) for s in st), 0, p % 16 == 0)
        order = [None] * len(groups)
        for (p, _), g in groups.items():
            order[g] = p
        for k in range(len(lay)):
            if lay[k] is None:
                lay[k] = cg.ExtLayout(0, (0,), 0, True)
        return lay, order

    def kernel_for(self, dims, lay):
        key = (tuple(sorted(dims.items())), tuple((e.grp, e.strides, e.offset, e.aligned16) for e in lay))
        hit = self._kernels.get(key)
        if hit is None:
            box = {}

            def gen(kn):
                box["spec"] = cg.gen_region_kernel(self.plan, dims, lay, kn)
                return box["spec"].source

            fn, kname = jit.get_function_gen(gen, "ptk_rowfuse")
            spec = box["spec"]
            spec.name, spec.finish_name = kname, kname + "_fin"
            if _lib.TRACE_ONLY:
                fin, blocks = 0, 2
            else:
                fin = jit.get_function(spec.source.replace("PTKKERNELNAMEPLACEHOLDER", kname), spec.finish_name)
                if spec.smem_bytes > 48 * 1024:
                    _lib.check(_lib.lib().ptk_func_set_max_dynamic_smem(fn, spec.smem_bytes), "max dynamic smem")
                nb = c_int(0)
                _lib.check(_lib.lib().ptk_func_max_active_blocks(fn, cg.WARPS * 32, spec.smem_bytes, ctypes.byref(nb)),
                           "occupancy")
                blocks = max(1, int(nb.value))
            hit = (spec, fn, fin, blocks)
            if len(self._kernels) > 16:
                self._kernels.clear()
            self._kernels[key] = hit
        return hit

    def _fused(self, vals):
        plan = self.plan
        B, dims, tens = self._bind(vals)
        lay, ptr_order = self._layouts(tens)
        spec, fn, fin, blocks = self.kernel_for(dims, lay)
        outs = [None] * self.n_out
        for v in plan.vals:
            if v.out < 0:
                continue
            if v.kind == "R1":
                shp = (B, dims[v.dom])
            elif v.kind == "R0":
                shp = (B,) if v.nd == 1 else (B, 1)
            elif v.kind == "S1":
                shp = (dims[v.dom],) if v.nd == 1 else (1, dims[v.dom])
            else:
                shp = (1,) * v.nd
            outs[v.out] = dev.empty(shp, v.dtype)
        grid = int(min((B + cg.WARPS - 1) // cg.WARPS, _lib.sm_count() * blocks))
        partials = dev.empty((grid * max(spec.acc_len, 1),), "float64")
        flag = nodes_basic._err_flag(f"{self.name}: index out of bounds")
        args = [c_void_p(p) for p in ptr_order]
        args += [c_void_p(dev.ptr(outs[k])) for k in spec.out_order]
        args += [c_void_p(dev.ptr(partials)), c_void_p(flag), c_longlong(B)]
        sp = dev.stream_ptr()
        jit.launch(fn, (grid,), (cg.WARPS * 32,), jit.KernelArgs(args), spec.smem_bytes, sp)
        if spec.acc_len:
            fargs = [c_void_p(dev.ptr(partials)), c_int(grid)] + [c_void_p(dev.ptr(outs[oi])) for oi, _, _, _ in spec.csum_out]
            jit.launch(fin, ((spec.acc_len * 32 + 255) // 256,), (256,), jit.KernelArgs(fargs), 0, sp)
        self.fused_calls += 1
        return [Val(d=o) for o in outs]

    def run(self, vals):
        if os.environ.get("PTK_ROWFUSE") == "0":
            return self._unfused(vals)
        try:
            return self._fused(vals)
        except cg.NotFusable as e:
            self.last_reason = str(e)
            return self._unfused(vals)
