"""CPU suite for the row-fused region kernel (codegen/rowfuse.py): the region finder on the lowered cfg5 program, and the
GENERATED kernel executed on the host (tests/kernel_emulator.py: every simulated thread is an OS thread, shuffles /
__syncwarp / shared-memory atomics are real barriers and atomics) against the reference C linker."""

import ctypes
from ctypes import c_int, c_longlong, c_void_p

import numpy as np
import pytest

from helpers import pytensor

import pytensor.tensor as pt
from kernel_emulator import EmulatedKernel
from pytensor_b200 import workloads as W
from pytensor_b200.codegen import rowfuse as cg
from pytensor_b200.vm.nodes_elemwise import ElemwiseNode, host_eval_program
from pytensor_b200.vm.nodes_rowfuse import RowRegionNode
from pytensor_b200.vm.values import Val


class _NpT:
    """NumPy array with the tensor interface `RowRegionNode._layouts` needs (element strides)."""

    def __init__(self, a):
        self.a = a

    def stride(self, i):
        return self.a.strides[i] // self.a.itemsize


def _aligned_copy(a, align=64):
    raw = np.empty(a.nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
    out[...] = a
    return out


def _host_values_until(program, stop_idx, inputs):
    """Evaluate the (host-side, metadata-only) steps in front of the region node on NumPy values."""
    vals = {}
    for s, arr in program.constants.items():
        vals[s] = Val(h=np.asarray(arr))
    for s, x in zip(program.inputs, inputs):
        vals[s] = Val(h=np.asarray(x))
    for st in program.steps[:stop_idx]:
        if type(st.impl) is ElemwiseNode:
            res = [Val(h=r) for r in host_eval_program(st.impl.prog, [np.asarray(vals[j].h) for j in st.ins])]
        else:
            res = st.impl.run([vals[j] for j in st.ins])
        for j, r in zip(st.outs, res):
            vals[j] = r
    return vals


def run_region_on_host(node: RowRegionNode, ext_arrays, tmp_path, grid=3):
    """Generate the node's kernel for these operands, run it + the finishing kernel in the emulator; returns outputs."""
    vals = [Val(h=a) for a in ext_arrays]
    plan = node.plan
    # bind without touching a device: shapes/dtypes come from the host values
    dims, B = {}, None
    tens = [None] * plan.n_ext
    for v in plan.vals:
        if v.ext < 0 or v.const is not None:
            continue
        a = ext_arrays[v.ext]
        assert a.dtype.name == v.dtype
        if v.kind == "R1":
            B = a.shape[0]
            dims[v.dom] = a.shape[1]
        elif v.kind == "R0":
            B = a.shape[0]
        elif v.kind == "S1":
            dims[v.dom] = a.shape[-1]
        elif v.kind == "S2":
            dims[v.dom], dims[v.dom2] = a.shape
        tens[v.ext] = _NpT(a)
    lay, ptr_order = node._layouts(tens, ptr_of=lambda t: t.a.ctypes.data)
    spec = cg.gen_region_kernel(plan, dims, lay, "k_region")
    main = EmulatedKernel(spec.source, "k_region", tmp_path, threaded=True, dynamic_smem="ptk_smem", warp_shim=True)
    fin = EmulatedKernel(spec.source, "k_region_fin", tmp_path, threaded=True, dynamic_smem="ptk_smem", warp_shim=True)
    outs = [None] * node.n_out
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
        outs[v.out] = np.full(shp, np.nan, dtype=v.dtype)
    partials = np.full(grid * max(spec.acc_len, 1), np.nan)
    err = np.zeros(1, dtype=np.int32)
    args = [c_void_p(p) for p in ptr_order] + [c_void_p(outs[k].ctypes.data) for k in spec.out_order]
    args += [c_void_p(partials.ctypes.data), c_void_p(err.ctypes.data), c_longlong(B)]
    main.launch(grid, cg.WARPS * 32, args)
    if spec.acc_len:
        fargs = [c_void_p(partials.ctypes.data), c_int(grid)] + [c_void_p(outs[oi].ctypes.data) for oi, _, _, _ in spec.csum_out]
        fin.launch((spec.acc_len * 32 + 255) // 256, 256, fargs)
    return outs, int(err[0]), spec


def _region_of(f):
    steps = f.vm.executor.program.steps
    idx = [i for i, st in enumerate(steps) if isinstance(st.impl, RowRegionNode)]
    assert len(idx) == 1, [repr(s.impl) for s in steps]
    return idx[0], steps[idx[0]]


@pytest.mark.parametrize("B,n,J,K", [(72, 96, 16, 8), (65, 50, 7, 3)])
def test_cfg5_region_kernel_on_the_host_matches_the_c_linker(tmp_path, B, n, J, K):
    pytensor.config.floatX = "float32"
    ins, outs, mk, meta = W.cfg5_logp_grad(B=B, n=n, J=J, K=K, dtype="float32", packed=False)
    f = pytensor.function(ins, outs, mode="CUDA")
    f_ref = pytensor.function(ins, outs, mode="CVM")
    args = mk(seed=7)
    exp = f_ref(*args)
    prog = f.vm.executor.program
    ridx, rstep = _region_of(f)
    # everything that moves data is inside the region: what is left are views, shape integers and the final reshapes
    assert len(prog.steps) <= 12
    vals = _host_values_until(prog, ridx, [_aligned_copy(a) for a in args])
    ext = [np.asarray(vals[s].h) for s in rstep.ins]
    got, err, spec = run_region_on_host(rstep.impl, ext, tmp_path)
    assert err == 0
    assert spec.n_loops <= 6
    # the shared group-index vector is wrapped and checked once per CTA (int32 table in shared memory), then read by both
    # the gather and the scatter of every row
    assert spec.source.count("cix0[j]") == 2 and "cix1" not in spec.source
    by_slot = dict(zip(rstep.outs, got))
    # map node outputs to function outputs through the trailing view steps
    after = {}
    for st in prog.steps[ridx + 1:]:
        after[st.outs[0]] = st.ins[0] if len(st.ins) == 1 else None
    for k, slot in enumerate(prog.outputs):
        while slot in after and after[slot] is not None:
            slot = after[slot]
        g = np.asarray(by_slot[slot]).reshape(np.shape(exp[k]))
        scale = max(1.0, float(np.max(np.abs(exp[k]))))
        np.testing.assert_allclose(g, exp[k], rtol=2e-5, atol=2e-5 * scale)


def test_region_flags_an_out_of_bounds_index(tmp_path):
    pytensor.config.floatX = "float32"
    ins, outs, mk, meta = W.cfg5_logp_grad(B=64, n=40, J=5, K=4, dtype="float32", packed=False)
    f = pytensor.function(ins, outs, mode="CUDA")
    args = mk(seed=3)
    args[6] = args[6].copy()
    args[6][11] = 5   # == J: out of bounds
    prog = f.vm.executor.program
    ridx, rstep = _region_of(f)
    vals = _host_values_until(prog, ridx, [_aligned_copy(a) for a in args])
    _, err, _ = run_region_on_host(rstep.impl, [np.asarray(vals[s].h) for s in rstep.ins], tmp_path, grid=2)
    assert err == 1


def test_region_finder_leaves_unrelated_graphs_alone():
    pytensor.config.floatX = "float32"
    for build in (W.cfg2_fused_elemwise, lambda: W.cfg3_mlp(256), lambda: W.metric_graph(n=32, layers=3, scan_steps=4)):
        ins, outs, mk, _ = build() if build is not W.cfg2_fused_elemwise else build(256)
        f = pytensor.function(ins, outs, mode="CUDA")
        assert not any(isinstance(st.impl, RowRegionNode) for st in f.vm.executor.program.steps)


def test_gather_elemwise_reduce_region_with_row_outputs(tmp_path):
    """A smaller region with per-row OUTPUTS: out1[b, j] = exp(w[b, idx[j]] * s[b]) (B, n) and out2[b] = sum_j out1."""
    pytensor.config.floatX = "float32"
    w, s = pt.fmatrix("w"), pt.fvector("s")
    idx = pt.lvector("idx")
    o1 = pt.exp(w[:, idx] * s[:, None])
    o2 = o1.sum(axis=1)
    f = pytensor.function([w, s, idx], [o1, o2], mode="CUDA")
    f_ref = pytensor.function([w, s, idx], [o1, o2], mode="CVM")
    rng = np.random.default_rng(5)
    args = [rng.standard_normal((70, 9)).astype("float32"), rng.standard_normal(70).astype("float32") * 0.3,
            rng.integers(-9, 9, size=33).astype("int64")]
    exp = f_ref(*args)
    prog = f.vm.executor.program
    ridx, rstep = _region_of(f)
    vals = _host_values_until(prog, ridx, [_aligned_copy(a) for a in args])
    got, err, spec = run_region_on_host(rstep.impl, [np.asarray(vals[x].h) for x in rstep.ins], tmp_path, grid=2)
    assert err == 0
    by_slot = dict(zip(rstep.outs, got))
    for k, slot in enumerate(prog.outputs):
        np.testing.assert_allclose(np.asarray(by_slot[slot]).reshape(exp[k].shape), exp[k], rtol=2e-5, atol=1e-5)


def test_per_cta_index_table_and_per_use_checks_agree(tmp_path, monkeypatch):
    """PTK_ROWFUSE_CTA_INDEX=0 keeps the per-use 64-bit bounds checks: same results, negative indices wrapped in both."""
    pytensor.config.floatX = "float32"
    ins, outs, mk, meta = W.cfg5_logp_grad(B=40, n=70, J=9, K=4, dtype="float32", packed=False)
    args = mk(seed=11)
    args[6] = args[6].copy()
    args[6][::3] -= 9     # NumPy-style negative group indices
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PTK_ROWFUSE_CTA_INDEX", flag)
        f = pytensor.function(ins, outs, mode="CUDA")
        prog = f.vm.executor.program
        ridx, rstep = _region_of(f)
        vals = _host_values_until(prog, ridx, [_aligned_copy(a) for a in args])
        (tmp_path / flag).mkdir()
        got, err, spec = run_region_on_host(rstep.impl, [np.asarray(vals[s].h) for s in rstep.ins], tmp_path / flag, grid=2)
        assert err == 0 and ("cix0" in spec.source) == (flag == "1")
        res.append(got)
    for a, b in zip(*res):   # (the shared-memory float atomics add in thread-arrival order: equal up to that rounding)
        a, b = np.asarray(a), np.asarray(b)
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=2e-5 * max(1.0, float(np.max(np.abs(b))) if b.size else 1.0))


def test_a_region_that_reads_a_view_of_another_regions_output_stays_behind_it():
    """Two regions where the second reads (through a transposing view) what the first produces: the pass used to emit the
    reader's fused node before its producer, and the reader's steps once more behind it (found by a random-graph search
    that interprets every lowered program with the NumPy port oracle).  Now the conflicting region stays unfused; the
    program is well ordered and agrees with the C linker."""
    from oracle import numpy_port
    from pytensor_b200.link.cuda.fusion import read_before_write

    pytensor.config.floatX = "float64"
    b, idx = pt.dmatrix("b"), pt.lvector("idx")
    h = pt.switch(b > 0, b, pt.expm1(b))
    u = pt.inc_subtensor(h[:, idx], h[:, idx] * 2)            # region over the rows of h
    ht = h.T
    w = pt.inc_subtensor(ht[:, idx], ht[:, idx] * 2)          # region over the rows of h.T: reads a view of the first one's output
    outs = [ht.sum(), w, u]
    rng = np.random.default_rng(1158)
    vals = [rng.standard_normal((6, 6)), np.array([1, -2, 0, 1])]
    f = pytensor.function([b, idx], outs, mode="CUDA")
    prog = f.vm.executor.program
    assert read_before_write(prog.steps, set(prog.inputs) | set(prog.constants)) is None
    exp = pytensor.function([b, idx], outs, mode="CVM")(*vals)
    got = numpy_port.evaluate_program(prog, vals)
    for g, e in zip(got, exp):
        np.testing.assert_allclose(g, e, rtol=1e-12, atol=1e-12)
