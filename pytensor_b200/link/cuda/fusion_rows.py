"""Region finder for the row-fused kernel (codegen/rowfuse.py, vm/nodes_rowfuse.py): SURVEY.md §8(f).4, the re-fusion
"beyond the reference's boundaries" that VERDICT r1 asks for.

Works on the lowered step list.  Starting from gathers / scatter-adds along axis 1 with a shared index vector
(`x[:, idx]`, AdvancedSubtensor / AdvancedIncSubtensor — the signature of a vectorised hierarchical model), it grows a
set of steps whose tensors all have the batch on axis 0 and are independent across it: Elemwise, row reductions, the
final sums over the batch, matrix products against a shared matrix (`Gemm` / `Dot22`; whether one side is skinny enough
is only known at run time, where the node falls back to its constituent steps otherwise), (B,) <-> (B,1) reshapes.
Every tensor gets a kind (R1 per-row vector, R0 per-row scalar, S1 / S0 / S2 shared) by propagation from the static
types (ndim + broadcast pattern).  The reference's own direction: pytensor/tensor/rewriting/fused_elemwise.py:107,278.
"""

from __future__ import annotations

import os

import numpy as np

from pytensor_b200.codegen.rowfuse import ROp, RV, RegionPlan
from pytensor_b200.vm import nodes_basic as nb
from pytensor_b200.vm import nodes_blas as nblas
from pytensor_b200.vm.nodes_elemwise import CAReduceNode, ElemwiseNode
from pytensor_b200.vm.nodes_rowfuse import RowRegionNode
from pytensor_b200.vm.vm import Step

_FLOATS = ("float32", "float64")
_INTS = ("int8", "int16", "int32", "int64", "uint8", "uint16", "uint32")
_EW_KIND = {(False, False): "R1", (False, True): "R0", (True, False): "S1", (True, True): "S0"}


class _Conflict(Exception):
    pass


def _anchor(st, types):
    impl = st.impl
    if type(impl) is nb.TakeNode and impl.axis == 1 and impl.naxes == 1 and len(st.ins) == 2:
        x, idx = (types.get(s) for s in st.ins)
        return bool(x and idx and x[1] == 2 and idx[1] == 1 and idx[0] in _INTS and x[0] in _FLOATS)
    if (type(impl) is nb.PutNode and impl.axis == 1 and impl.naxes == 1 and not impl.set_instead_of_inc
            and len(st.ins) == 3):
        x, y, idx = (types.get(s) for s in st.ins)
        return bool(x and y and idx and x[1] == 2 and y[1] == 2 and idx[1] == 1 and idx[0] in _INTS
                    and x[0] in _FLOATS and y[0] == x[0] and impl.dtype == x[0])
    return False


def _classify(st, types, kinds):
    """Kinds this step implies for its slots ({slot: kind}) if it is a fusable op form, else None.  Unknown kinds are
    inferred from the known ones; raises _Conflict when the known ones contradict the op."""
    impl = st.impl
    t = type(impl)
    ins, outs = st.ins, st.outs
    want = {}

    def put(slot, kind):
        if want.setdefault(slot, kind) != kind:
            raise _Conflict

    if t is ElemwiseNode:
        if impl.ndim == 2:
            ok = types.get(outs[0])
            if ok is None:
                return None
            okind = _EW_KIND[tuple(ok[2])]
            if okind not in ("R1", "R0"):
                return None
            for s, bc in zip(ins, impl.in_bcast):
                put(s, _EW_KIND[tuple(bc)])
            for s in outs:
                put(s, okind)
            return want
        if impl.ndim == 1:
            vec = [s for s, bc in zip(ins, impl.in_bcast) if not bc[0]] + list(outs)
            known = {kinds[s] for s in vec if s in kinds}
            if known != {"R0"}:
                return None
            for s, bc in zip(ins, impl.in_bcast):
                put(s, "S0" if bc[0] else "R0")
            for s in outs:
                put(s, "R0")
            return want
        return None
    if t is CAReduceNode:
        k = kinds.get(ins[0])
        if impl.in_dtype not in _FLOATS or impl.out_dtype not in _FLOATS:
            return None
        if k == "R1" and impl.ndim == 2:
            if impl.axes == (1,) and impl.red_op in ("add", "mul", "maximum", "minimum"):
                put(ins[0], "R1"), put(outs[0], "R0")
                return want
            if impl.axes in ((0,), (0, 1)) and impl.red_op == "add":
                put(ins[0], "R1"), put(outs[0], "S1" if impl.axes == (0,) else "S0")
                return want
        if k == "R0" and impl.ndim == 1 and impl.axes == (0,) and impl.red_op == "add":
            put(ins[0], "R0"), put(outs[0], "S0")
            return want
        return None
    if t is nb.DimShuffleNode:
        ki, ko = kinds.get(ins[0]), kinds.get(outs[0])
        if "R0" in (ki, ko) and ((impl.input_ndim == 1 and impl.new_order == (0, "x"))
                                 or (impl.input_ndim == 2 and impl.new_order == (0,))):
            put(ins[0], "R0"), put(outs[0], "R0")
            return want
        return None
    if t in (nblas.GemmNode, nblas.Dot22Node):
        if t is nblas.GemmNode:
            z, a, x, y, b = ins
        else:
            if impl.scalar:
                return None
            z, (x, y) = None, ins
        if impl.dtype not in _FLOATS or getattr(impl, "precision", 0) != 0:
            return None
        if not any(kinds.get(s) == "R1" for s in ([x, outs[0]] + ([z] if z is not None else []))):
            return None
        put(x, "R1"), put(y, "S2"), put(outs[0], "R1")
        if z is not None:
            zt = types.get(z)
            if zt is None or tuple(zt[2]) != (False, False):
                return None
            put(z, "R1"), put(a, "S0"), put(b, "S0")
        return want
    if t is nb.TakeNode and _anchor(st, types):
        put(ins[0], "R1"), put(ins[1], "S1"), put(outs[0], "R1")
        return want
    if t is nb.PutNode and _anchor(st, types):
        put(ins[0], "R1"), put(ins[1], "R1"), put(ins[2], "S1"), put(outs[0], "R1")
        return want
    return None


def fuse_row_regions(steps, output_slots, opts):
    types = opts.get("slot_types")
    if not types or os.environ.get("PTK_ROWFUSE_PASS") == "0" or not any(_anchor(st, types) for st in steps):
        return steps
    consts = opts.get("constants", {})
    kinds = {}
    member = set()

    def try_join(i):
        st = steps[i]
        try:
            want = _classify(st, types, kinds)
        except _Conflict:
            return False
        if want is None:
            return False
        for s, k in want.items():
            if s in kinds and kinds[s] != k:
                return False
        connected = _anchor(st, types) or any(kinds.get(s) in ("R1", "R0") for s in want)
        if not connected:
            return False
        for s, k in want.items():
            tp = types.get(s)
            if tp is None:
                return False
            if k in ("R1", "R0") and tp[0] not in _FLOATS:
                return False
            if k == "S2" and tp[0] not in _FLOATS:
                return False
        kinds.update(want)
        member.add(i)
        return True

    changed = True
    while changed:
        changed = False
        for i in range(len(steps)):
            if i not in member and try_join(i):
                changed = True
    if not member:
        return steps

    # connected components over R-kind slots
    parent = {i: i for i in member}

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    owner = {}
    for i in sorted(member):
        st = steps[i]
        for s in list(st.ins) + list(st.outs):
            if kinds.get(s) in ("R1", "R0"):
                if s in owner:
                    parent[find(i)] = find(owner[s])
                else:
                    owner[s] = i
    comps = {}
    for i in member:
        comps.setdefault(find(i), []).append(i)

    new_steps = list(steps)
    removed = set()
    inserts = {}     # position (index of the region's last step) -> (fused Step, tainted outside steps to move behind it)
    # Regions are accepted in program order of their last step, and only while the steps they TOUCH (members + outside steps
    # that move behind the fused node) are untouched by an accepted region: a region whose members another region has
    # moved (it reads, through a view, what that region produces) would otherwise be emitted BEFORE its producer — and
    # its constituent steps once more behind it.  The rejected region simply stays unfused.
    for comp in sorted((sorted(c) for c in comps.values()), key=lambda c: c[-1]):
        built = _build_region(steps, comp, kinds, types, consts, set(output_slots), opts)
        if built is None:
            continue
        fused_step, moved = built
        touch = set(comp) | set(moved)
        if touch & removed:
            continue
        removed.update(touch)
        inserts[comp[-1]] = (fused_step, [steps[j] for j in moved])
    if not inserts:
        return steps
    out = []
    for i, st in enumerate(new_steps):
        if i in inserts:
            fs, moved = inserts[i]
            out.append(fs)
            out.extend(moved)
        elif i not in removed:
            out.append(st)
    return out


def _build_region(steps, comp, kinds, types, consts, outset, opts):
    comp_set = set(comp)
    produced = {}
    for i in comp:
        for s in steps[i].outs:
            produced[s] = i
    if not any(_anchor(steps[i], types) for i in comp):
        return None
    # S-kind operands must come from outside (cross-row results are terminal inside one kernel)
    for i in comp:
        for s in steps[i].ins:
            if kinds.get(s) in ("S0", "S1", "S2") and s in produced:
                return None
    # steps between the region's first and last step that depend on the region move behind the fused node; the region
    # must not depend on them, and nothing that stays in front may write into its inputs in place
    first, last = comp[0], comp[-1]
    tainted_slots = set(produced)
    moved = []
    for j in range(first, last + 1):
        if j in comp_set:
            if any(s in tainted_slots and s not in produced for s in steps[j].ins):
                return None
            continue
        st = steps[j]
        if any(s in tainted_slots for s in st.ins):
            moved.append(j)
            tainted_slots.update(st.outs)
        elif getattr(st.impl, "destroy", None):
            return None
    # values
    vals, vid = [], {}
    alias = {}

    def canon(s):
        while s in alias:
            s = alias[s]
        return s

    # dimension symbols by unification
    sym_parent = {}

    def sfind(x):
        while sym_parent.setdefault(x, x) != x:
            sym_parent[x] = sym_parent[sym_parent[x]]
            x = sym_parent[x]
        return x

    def sunion(a, b):
        ra, rb = sfind(a), sfind(b)
        if ra != rb:
            sym_parent[ra] = rb

    for i in comp:
        st = steps[i]
        if type(st.impl) is nb.DimShuffleNode:
            alias[st.outs[0]] = st.ins[0]

    def dimsym(s, axis):
        return ("d", canon(s), axis)

    ext_slots, out_slots = [], []
    baked = {}
    readers_outside = set()
    for j, st in enumerate(steps):
        if j not in comp_set:
            readers_outside.update(st.ins)

    def value(s):
        c = canon(s)
        if c in vid:
            return vid[c]
        k = kinds[c] if c in kinds else kinds[s]
        tp = types[c]
        rv = RV(k, tp[0], nd=tp[1])
        if c not in produced:
            if k == "S0" and c in consts and np.asarray(consts[c]).size == 1:
                rv.const = np.asarray(consts[c]).reshape(-1)[0].item()
                baked[c] = np.asarray(consts[c])
            else:
                rv.ext = len(ext_slots)
                ext_slots.append(c)
        vals.append(rv)
        vid[c] = len(vals) - 1
        return vid[c]

    ops = []
    for i in comp:
        st = steps[i]
        impl = st.impl
        t = type(impl)
        if t is nb.DimShuffleNode:
            value(st.ins[0])
            continue
        if t is ElemwiseNode:
            iv = [value(s) for s in st.ins]
            ov = [value(s) for s in st.outs]
            if vals[ov[0]].kind == "R1":
                for s, v in zip(st.ins, iv):
                    if vals[v].kind in ("R1", "S1"):
                        sunion(dimsym(s, 1 if vals[v].kind == "R1" or types[canon(s)][1] == 2 else 0), dimsym(st.outs[0], 1))
                for s in st.outs[1:]:
                    sunion(dimsym(s, 1), dimsym(st.outs[0], 1))
            elif any(vals[v].kind in ("R1", "S1") for v in iv):
                return None
            ops.append(ROp("ew", iv, ov, prog=impl.prog, name=impl.name))
        elif t is CAReduceNode:
            iv, ov = value(st.ins[0]), value(st.outs[0])
            kin, kout = vals[iv].kind, vals[ov].kind
            if kin == "R1" and kout == "R0":
                ops.append(ROp("rsum", [iv], [ov], red=impl.red_op, acc_dtype=impl.acc_dtype, name=impl.name))
            elif kin == "R1" and kout == "S1":
                sunion(dimsym(st.outs[0], 0), dimsym(st.ins[0], 1))
                ops.append(ROp("csum", [iv], [ov], acc_dtype=impl.acc_dtype, name=impl.name))
            elif kin == "R1" and kout == "S0":   # total: row sums (kept in the accumulator type) then the batch sum
                vals.append(RV("R0", impl.acc_dtype if impl.acc_dtype in _FLOATS else "float64"))
                mid = len(vals) - 1
                ops.append(ROp("rsum", [iv], [mid], red="add", acc_dtype=vals[mid].dtype, name=impl.name))
                ops.append(ROp("csum", [mid], [ov], acc_dtype=impl.acc_dtype, name=impl.name))
            else:
                ops.append(ROp("csum", [iv], [ov], acc_dtype=impl.acc_dtype, name=impl.name))
            if impl.acc_dtype not in _FLOATS:
                return None
        elif t in (nblas.GemmNode, nblas.Dot22Node):
            if t is nblas.GemmNode:
                z, a, x, y, b = st.ins
                zv, av, bv = value(z), value(a), value(b)
                sunion(dimsym(z, 1), dimsym(st.outs[0], 1))
            else:
                x, y = st.ins
                zv = -1
                vals.append(RV("S0", impl.dtype, const=1.0, nd=0))
                av = len(vals) - 1
                vals.append(RV("S0", impl.dtype, const=0.0, nd=0))
                bv = len(vals) - 1
            xv, yv, ov = value(x), value(y), value(st.outs[0])
            sunion(dimsym(x, 1), dimsym(y, 0))
            sunion(dimsym(st.outs[0], 1), dimsym(y, 1))
            ops.append(ROp("gemm", [zv, xv, yv], [ov], alpha=av, beta=bv, name=impl.name))
        elif t is nb.TakeNode:
            xv, iv, ov = value(st.ins[0]), value(st.ins[1]), value(st.outs[0])
            sunion(dimsym(st.outs[0], 1), dimsym(st.ins[1], 0))
            ops.append(ROp("take", [xv, iv], [ov], name=impl.name))
        elif t is nb.PutNode:
            xv, yv, iv, ov = value(st.ins[0]), value(st.ins[1]), value(st.ins[2]), value(st.outs[0])
            sunion(dimsym(st.ins[1], 1), dimsym(st.ins[2], 0))
            sunion(dimsym(st.outs[0], 1), dimsym(st.ins[0], 1))
            ops.append(ROp("put", [xv, yv, iv], [ov], name=impl.name))
        else:
            return None
    # outputs: region-produced slots that something outside reads (or that are program outputs)
    seen_out = set()
    for i in comp:
        for s in steps[i].outs:
            if s in readers_outside or s in outset:
                c = canon(s)
                if c not in produced:
                    return None      # a reshaped view of a region INPUT escapes: leave such graphs alone
                v = vid.get(c)
                if v is None or v in seen_out:
                    return None      # two escaping views of one value
                seen_out.add(v)
                vals[v].out = len(out_slots)
                vals[v].nd = types[s][1]
                out_slots.append(s)
    if not out_slots:
        return None
    # domain symbols -> small integers
    symid = {}

    def sid(key):
        r = sfind(key)
        return symid.setdefault(r, len(symid))

    for c, v in vid.items():
        rv = vals[v]
        nd = types[c][1]
        if rv.kind == "R1":
            rv.dom = sid(dimsym(c, 1))
        elif rv.kind == "S1":
            rv.dom = sid(dimsym(c, 1 if nd == 2 else 0))
        elif rv.kind == "S2":
            rv.dom, rv.dom2 = sid(dimsym(c, 0)), sid(dimsym(c, 1))
    for rv in vals:  # internal helper values created above carry no slot: R0 only
        if rv.kind in ("R1", "S1") and rv.dom < 0:
            return None
    plan = RegionPlan(vals, ops, n_ext=len(ext_slots), n_out=len(out_slots))
    # the unfused path: the region's own steps with the usual peepholes applied
    from pytensor_b200.link.cuda.fusion_passes import fuse_elemwise_reduce, fuse_gemm_epilogue

    sub = [steps[i] for i in comp]
    sub = fuse_elemwise_reduce(fuse_gemm_epilogue(sub, out_slots, opts), out_slots, opts)
    # in-place writes of the unfused path into the node's inputs, for the VM's hazard tracking
    destroy = {}
    ext_index = {s: k for k, s in enumerate(ext_slots)}
    for stp in sub:
        for o, k in (getattr(stp.impl, "destroy", None) or {}).items():
            if k < len(stp.ins) and stp.ins[k] in ext_index:
                if len(destroy) >= len(out_slots):
                    return None
                destroy[len(destroy)] = ext_index[stp.ins[k]]
    names = [type(steps[i].impl).__name__.replace("Node", "") for i in comp]
    node = RowRegionNode(plan, sub, ext_slots, out_slots, destroy,
                         name=f"RowRegion[{len(comp)} steps: " + ",".join(sorted(set(names))) + "]", consts=baked)
    origin = steps[comp[-1]].origin
    return Step(node, ext_slots, out_slots, origin=origin), moved
