"""CPU suite: the scalar-op expression table (pytensor_b200/codegen/scalar.py) against the reference C linker, op by op.

The generated scalar bodies are plain C++.  Compiled for the host with a small shim for the few device intrinsics, they
are run over vectors of test values and compared with the same graph under mode="CVM" — i.e. with the C expressions of
`ScalarOp.c_code` (pytensor/scalar/basic.py:1411-3861, pytensor/scalar/math.py) and the same libm.  This validates the
STRUCTURE of every expression (operand order, casts, floor-division / modulo sign rules, NaN propagation of max/min,
Softplus / Log1mexp branch thresholds, integer true division through double, bool arithmetic) without a GPU; the accuracy
of CUDA's own libm is what the -m gpu parity tests measure.  Ops that need CUDA-only math functions (erfcx, erfinv,
erfcinv, cyl_bessel_i0/i1) are left to the GPU suite."""

import ctypes
import subprocess

import numpy as np
import pytest

from helpers import pytensor

import pytensor.tensor as pt
from pytensor_b200.codegen.scalar import CTYPE, emit_body

SHIM = r"""
#include <cmath>
#include <cstring>
#include <cstdint>
#define __device__
#define __forceinline__ inline
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline double __longlong_as_double(long long v) { double f; std::memcpy(&f, &v, 8); return f; }
static inline float ptk_max_nan_f32(float a, float b) { return (b > a) ? b : ((a >= b) ? a : NAN); }
static inline float ptk_min_nan_f32(float a, float b) { return (b < a) ? b : ((a <= b) ? a : NAN); }
using std::isnan; using std::isinf;
template <typename T> static inline T ptk_floordiv(T x, T y) { if (y == 0) return 0; T q = x / y; if ((x % y != 0) && ((x < 0) != (y < 0))) --q; return q; }
template <typename T> static inline T ptk_imod_py(T x, T y) { if (y == 0) return 0; T r = x % y; if (r != 0 && ((r < 0) != (y < 0))) r += y; return r; }
template <typename T> static inline T ptk_fmod_py(T x, T y) { T r = std::fmod(x, y); if (r != 0 && ((r < 0) != (y < 0))) r += y; return r; }
"""


def _compile_body(prog, tmp_path, tag, extra_flags=()):
    ins_decl = ", ".join(f"const {CTYPE[d]}* i{k}" for k, d in enumerate(prog.in_dtypes))
    outs_decl = ", ".join(f"{CTYPE[d]}* o{k}" for k, d in enumerate(prog.out_dtypes))
    call = ", ".join([f"i{k}[n]" for k in range(len(prog.in_dtypes))] + [f"o{k}[n]" for k in range(len(prog.out_dtypes))])
    src = SHIM + emit_body(prog) + (f'\nextern "C" void run(long long N, {ins_decl}, {outs_decl}) '
                                    f'{{ for (long long n = 0; n < N; ++n) ptk_body({call}); }}\n')
    cpp, so = tmp_path / f"body{tag}.cpp", tmp_path / f"body{tag}.so"
    cpp.write_text(src)
    # same floating-point environment as the reference's generated C (`-march=native`, GNU default FMA contraction,
    # pytensor/link/c/cmodule.py) — nvcc contracts mul+add into FMA as well; which pairs get fused may still differ,
    # hence the small absolute tolerance on cancelling expressions below
    subprocess.run(["g++", "-O2", "-march=native", "-fno-math-errno", *extra_flags, "-shared", "-fPIC", "-std=c++17", str(cpp),
                    "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def _emulate(inputs, outputs, values, tmp_path, extra_flags=()):
    """Run the lowered program on the host: every step must be a plain ElemwiseNode over same-shaped vectors."""
    f = pytensor.function(inputs, outputs, mode="CUDA")
    prog = f.vm.executor.program
    slots = dict(zip(prog.inputs, [np.ascontiguousarray(v) for v in values]))
    for s, c in prog.constants.items():
        slots[s] = np.asarray(c)
    N = len(values[0])
    for k, st in enumerate(prog.steps):
        assert type(st.impl).__name__ == "ElemwiseNode", f"step {k} is {type(st.impl).__name__}: keep the test graph elementwise"
        p = st.impl.prog
        lib = _compile_body(p, tmp_path, k, extra_flags)
        ins = []
        for j, d in zip(st.ins, p.in_dtypes):
            a = np.ascontiguousarray(np.broadcast_to(slots[j], (N,)).astype(d, copy=False))
            assert a.dtype == np.dtype(d)
            ins.append(a)
        outs = [np.empty(N, dtype=d) for d in p.out_dtypes]
        lib.run(ctypes.c_longlong(N), *[a.ctypes.data_as(ctypes.c_void_p) for a in ins],
                *[a.ctypes.data_as(ctypes.c_void_p) for a in outs])
        for j, o in zip(st.outs, outs):
            slots[j] = o
    got = [slots[s] for s in prog.outputs]
    ref = pytensor.function(inputs, outputs, mode="CVM")(*values)
    return got, ref


def _check(got, ref, rtol, atol=0.0):
    assert len(got) == len(ref)
    for k, (g, r) in enumerate(zip(got, ref)):
        r = np.asarray(r)
        g = g.view(np.bool_) if r.dtype == np.bool_ else g
        assert g.dtype == r.dtype, (k, g.dtype, r.dtype)
        if r.dtype.kind in "biu":
            np.testing.assert_array_equal(g, r, err_msg=f"output {k}")
        else:
            np.testing.assert_allclose(g, r, rtol=rtol, atol=atol, equal_nan=True, err_msg=f"output {k}")


def _floats(dtype, n=3000, scale=3.0, seed=7):
    rng = np.random.default_rng(seed)
    v = (rng.standard_normal(n) * scale).astype(dtype)
    v[:12] = np.array([0.0, -0.0, 1.0, -1.0, 0.5, -0.5, 2.5, -2.5, 1.5, -1.5, 40.0, -40.0], dtype=dtype)
    return v


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_unary_float_ops(tmp_path, dtype):
    pytensor.config.floatX = dtype
    x = pt.vector("x", dtype=dtype)
    pos = pt.abs(x) + 0.25
    unit = pt.clip(x * 0.3, -0.95, 0.95)
    outs = [pt.exp(x), pt.exp2(x), pt.expm1(x), pt.log(pos), pt.log2(pos), pt.log10(pos), pt.log1p(pt.abs(x)), pt.sqrt(pos),
            pt.sin(x), pt.cos(x), pt.tan(unit), pt.arcsin(unit), pt.arccos(unit), pt.arctan(x), pt.sinh(x), pt.cosh(x),
            pt.tanh(x), pt.arcsinh(x), pt.arccosh(pos + 1), pt.arctanh(unit), pt.erf(x), pt.erfc(x), pt.gamma(pos),
            pt.gammaln(pos), pt.sigmoid(x), pt.softplus(x * 12), pt.log1mexp(-pos), pt.abs(x), -x, pt.sign(x), pt.ceil(x),
            pt.floor(x), pt.trunc(x), pt.round(x, mode="half_to_even"), pt.round(x, mode="half_away_from_zero"),
            pt.reciprocal(pos), pt.sqr(x), pt.isnan(pt.log(x)), pt.isinf(pt.exp(x * 30))]
    got, ref = _emulate([x], outs, [_floats(dtype)], tmp_path)
    _check(got, ref, rtol=2e-6 if dtype == "float32" else 1e-13, atol=1e-30)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_binary_float_ops_nan_propagation_and_python_modulo(tmp_path, dtype):
    pytensor.config.floatX = dtype
    x, y = pt.vector("x", dtype=dtype), pt.vector("y", dtype=dtype)
    xv, yv = _floats(dtype, seed=8), _floats(dtype, seed=9)[::-1].copy()
    yv[yv == 0] = 1.5
    xv[20:24] = np.nan  # Maximum / Minimum propagate NaN like the reference's C (scalar/basic.py:2009-2062)
    yv[22:26] = np.nan
    outs = [x + y, x - y, x * y, x / y, x // y, x % y, pt.pow(pt.abs(x) + 0.1, y * 0.2), pt.maximum(x, y), pt.minimum(x, y),
            pt.arctan2(x, y), pt.eq(x, y), pt.neq(x, y), pt.lt(x, y), pt.le(x, y), pt.gt(x, y), pt.ge(x, y),
            pt.switch(x > y, x, y * 2), pt.clip(x, -1.0, y * 0 + 1.25), pt.true_div(x, y) + pt.sqr(y)]
    got, ref = _emulate([x, y], outs, [xv, yv], tmp_path)
    _check(got, ref, rtol=2e-6 if dtype == "float32" else 1e-13, atol=2e-6 if dtype == "float32" else 1e-14)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_special_functions_psi_trigamma_gammainc_betainc(tmp_path, dtype):
    """Psi / TriGamma / GammaInc / GammaIncC / BetaInc (scalar/math.py:403,502,627,674,1342) — the device restatements
    in codegen/scalar.py::SPECIAL_HELPERS, compiled for the host, against the reference's own C (gamma.c, incbet.c) over
    the regimes each algorithm switches between, plus the argument checks and limits the reference's tests pin
    (tests/scalar/test_math.py:34-80: nan and inf arguments)."""
    pytensor.config.floatX = dtype
    rng = np.random.default_rng(21)
    n = 4000
    x = pt.vector("x", dtype=dtype)
    xv = np.concatenate([rng.uniform(-30, 60, n - 16), [0.0, -1.0, -2.0, -0.5, -1.5, 1e-6, 1e-5, 9e-5, 1e-4, 8.5, 5.0, 1.0,
                                                       0.5, 100.0, 1e4, -1e-3]]).astype(dtype)
    got, ref = _emulate([x], [pt.psi(x), pt.tri_gamma(x)], [xv], tmp_path)
    _check(got, ref, rtol=1e-5 if dtype == "float32" else 1e-9, atol=1e-6 if dtype == "float32" else 1e-12)

    k, y = pt.vector("k", dtype=dtype), pt.vector("y", dtype=dtype)
    kv = np.concatenate([rng.uniform(0.05, 40, n - 12), [1, 1, np.inf, 1, np.inf, -1, 1, 0.5, 170, 171.5, 300, 1e-3]]).astype(dtype)
    yv = np.concatenate([rng.uniform(0, 80, n - 12), [2, np.inf, 1, np.nan, np.inf, 1, -1, 0, 160, 180, 310, 1e-3]]).astype(dtype)
    (tmp_path / "g").mkdir()  # a fresh directory per graph: dlopen caches handles by path
    got, ref = _emulate([k, y], [pt.gammainc(k, y), pt.gammaincc(k, y)], [kv, yv], tmp_path / "g")
    _check(got, ref, rtol=1e-5 if dtype == "float32" else 1e-9, atol=1e-7 if dtype == "float32" else 1e-13)

    a, b, z = pt.vector("a", dtype=dtype), pt.vector("b", dtype=dtype), pt.vector("z", dtype=dtype)
    av = np.concatenate([rng.uniform(0.05, 30, n - 10), [1, 2, 0.5, 100, 200, 1e-2, 5, -1, 3, 3]]).astype(dtype)
    bv = np.concatenate([rng.uniform(0.05, 30, n - 10), [1, 3, 0.5, 100, 150, 50, 1e-2, 2, -2, 3]]).astype(dtype)
    zv = np.concatenate([rng.uniform(0, 1, n - 10), [0.3, 0.0, 1.0, 0.5, 0.6, 0.01, 0.99, 0.5, 0.5, 1.5]]).astype(dtype)
    (tmp_path / "b").mkdir()
    got, ref = _emulate([a, b, z], [pt.betainc(a, b, z)], [av, bv, zv], tmp_path / "b")
    _check(got, ref, rtol=1e-5 if dtype == "float32" else 1e-9, atol=1e-7 if dtype == "float32" else 1e-13)


@pytest.mark.parametrize("dtype", ["int8", "int32", "int64", "uint8", "uint32"])
def test_integer_ops_floor_division_modulo_bitwise_and_true_division(tmp_path, dtype):
    pytensor.config.floatX = "float64"
    i, j = pt.vector("i", dtype=dtype), pt.vector("j", dtype=dtype)
    info = np.iinfo(dtype)
    rng = np.random.default_rng(10)
    lo, hi = max(info.min, -50), min(info.max, 50)
    iv = rng.integers(lo, hi + 1, size=2000).astype(dtype)
    jv = rng.integers(lo, hi + 1, size=2000).astype(dtype)
    jv[jv == 0] = 3  # (division by zero is undefined behaviour in the reference's C as well)
    outs = [i + j, i - j, i * j, i // j, i % j, i / j, i & j, i | j, i ^ j, ~i, pt.maximum(i, j), pt.minimum(i, j), pt.eq(i, j),
            pt.lt(i, j), pt.ge(i, j), pt.abs(i), pt.switch(i > j, i, j), pt.cast(i, "float32") * 0.5, pt.cast(i, "int16") + 1,
            pt.sqr(i), pt.clip(i, lo // 2 if lo < 0 else 1, hi // 2)]
    if info.min < 0:
        outs += [-i, pt.sign(i)]
    got, ref = _emulate([i, j], outs, [iv, jv], tmp_path)
    _check(got, ref, rtol=1e-13)


def test_bool_ops_and_mixed_casts(tmp_path):
    pytensor.config.floatX = "float64"
    a, b = pt.vector("a", dtype="bool"), pt.vector("b", dtype="bool")
    x = pt.dvector("x")
    rng = np.random.default_rng(11)
    av, bv = rng.integers(0, 2, size=500).astype(bool), rng.integers(0, 2, size=500).astype(bool)
    xv = rng.standard_normal(500) * 300
    outs = [a & b, a | b, a ^ b, ~a, pt.cast(a, "int32") + pt.cast(b, "int32"), pt.switch(a, x, -x), pt.cast(x, "int32"),
            pt.cast(x, "int8"), pt.cast(pt.abs(x), "uint8"), pt.cast(x, "float32"), pt.cast(x > 0, "float64") * x, pt.eq(a, b)]
    got, ref = _emulate([a, b, x], outs, [av, bv, xv], tmp_path)
    _check(got, ref, rtol=1e-13)


@pytest.mark.parametrize("dtype", ["int64", "int32", "uint8"])
def test_host_evaluator_for_shape_arithmetic_matches_reference(dtype):
    # integer Elemwise over <= 64-element HOST values (shape plumbing) is evaluated by `host_eval_program` instead of a
    # kernel (vm/nodes_elemwise.py); it must agree with the reference on the same integer graphs
    from pytensor_b200.vm.nodes_elemwise import host_eval_program

    i, j = pt.vector("i", dtype=dtype), pt.vector("j", dtype=dtype)
    info = np.iinfo(dtype)
    rng = np.random.default_rng(12)
    lo, hi = max(info.min, -50), min(info.max, 50)
    iv = rng.integers(lo, hi + 1, size=48).astype(dtype)
    jv = rng.integers(lo, hi + 1, size=48).astype(dtype)
    jv[jv == 0] = 3
    outs = [i + j * 2, i - j, i // j, i % j, pt.maximum(i, j), pt.minimum(i, j), pt.switch(i > j, i, j), pt.abs(i - j),
            pt.eq(i, j), pt.le(i, j), i & j, i | j, i ^ j, pt.sqr(i), pt.cast(i, "int64") * 3]
    f = pytensor.function([i, j], outs, mode="CUDA")
    prog = f.vm.executor.program
    slots = dict(zip(prog.inputs, [iv, jv]))
    for s, c in prog.constants.items():
        slots[s] = np.asarray(c)
    for st in prog.steps:
        assert type(st.impl).__name__ == "ElemwiseNode"
        res = host_eval_program(st.impl.prog, [np.broadcast_to(slots[k], iv.shape) for k in st.ins])
        assert res is not None, "every op of this graph belongs to the host evaluator's table"
        for k, r in zip(st.outs, res):
            slots[k] = np.asarray(r)
    got = [np.asarray(slots[s]) for s in prog.outputs]
    ref = pytensor.function([i, j], outs, mode="CVM")(iv, jv)
    _check(got, ref, rtol=0)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_max_of_a_value_and_its_negation_becomes_abs_bit_exactly(tmp_path, dtype, monkeypatch):
    """The lowering-time peephole Maximum(u, -u) -> Abs(u) (codegen/scalar.py::simplify): same BITS as the unsimplified
    program and as the reference C linker for NaNs, signed zeros, infinities, denormals and ordinary values, in both forms
    the host's canonicaliser produces (`Neg(u)`, and c*x next to (-c)*x); maxima that only LOOK similar are left alone."""
    from pytensor_b200.codegen import scalar as cs

    pytensor.config.floatX = dtype
    x, y = pt.vector("x", dtype=dtype), pt.vector("y", dtype=dtype)
    c = np.asarray(0.9, dtype=dtype)
    u = (x * y + np.asarray(0.5, dtype=dtype)) * c
    w = pt.exp(x)
    outs = [pt.maximum(u, -u) + pt.sqr(x) * np.asarray(0.1, dtype=dtype),     # c*t next to (-c)*t after canonicalisation
            pt.maximum(-w, w),                                                  # Neg form, operands swapped
            pt.maximum(x * c, y * (-c)),                                        # different operands: NOT an abs
            pt.maximum(x * c, x * np.asarray(-0.8, dtype=dtype))]               # different magnitudes: NOT an abs
    v = _floats(dtype, n=4000)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, np.finfo(dtype).tiny / 4, -np.finfo(dtype).tiny / 4,
                        np.finfo(dtype).max, -np.finfo(dtype).max], dtype=dtype)
    xs = np.concatenate([np.repeat(special, len(special)), v])
    ys = np.concatenate([np.tile(special, len(special)), v[::-1]])
    got, ref = _emulate([x, y], outs, [xs, ys], tmp_path)
    _check(got, ref, rtol=2e-6 if dtype == "float32" else 1e-13, atol=1e-30)
    f = pytensor.function([x, y], outs, mode="CUDA")
    ops = [i.op for st in f.vm.executor.program.steps for i in st.impl.prog.insts]
    assert ops.count("Abs") == 2 and ops.count("Maximum") == 2, ops
    # Against the unsimplified programs with FMA contraction off (contraction picks its pairs per expression tree and may
    # move a last bit in EITHER program, like in any other expression): identical bits.
    (tmp_path / "s").mkdir(), (tmp_path / "raw").mkdir()
    got_s, _ = _emulate([x, y], outs, [xs, ys], tmp_path / "s", extra_flags=("-ffp-contract=off",))
    monkeypatch.setenv("PTK_SCALAR_SIMPLIFY", "0")
    got_raw, _ = _emulate([x, y], outs, [xs, ys], tmp_path / "raw", extra_flags=("-ffp-contract=off",))
    f_raw = pytensor.function([x, y], outs, mode="CUDA")
    assert [i.op for st in f_raw.vm.executor.program.steps for i in st.impl.prog.insts].count("Abs") == 0
    uint = np.uint32 if dtype == "float32" else np.uint64
    for g, g0, r in zip(got_s, got_raw, ref):
        nan = np.isnan(g0)
        np.testing.assert_array_equal(np.isnan(g), nan)
        # (max(-0, +0): the reference's C expression — and this file's host stand-in — keeps its FIRST operand's zero, the
        # device's max.NaN.f32 returns +0 like |y| does; zeros therefore compare by value, everything else by bits)
        nz = ~nan & (g0 != 0)
        np.testing.assert_array_equal(g[nz].view(uint), g0[nz].view(uint))
        np.testing.assert_array_equal(g[~nan & ~nz], g0[~nan & ~nz])
        np.testing.assert_array_equal(np.isnan(np.asarray(r)), nan)                      # the C linker agrees on the NaNs
