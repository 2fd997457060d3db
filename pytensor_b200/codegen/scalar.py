"""Scalar program IR and its translation to CUDA device code.

A `ScalarProgram` is the backend's own, plain-data restatement of the straight-line scalar graph the reference keeps
in `Composite.fgraph` (pytensor/scalar/basic.py:4010-4170) or of a single `ScalarOp`.  The lowering
(pytensor_b200/link/cuda/lower.py) builds it; nothing in this module imports pytensor.

Semantics follow the per-op C expressions the reference's C linker compiles (SURVEY.md appendix C restates them from
`ScalarOp.c_code` in pytensor/scalar/basic.py:1411-3861 and pytensor/scalar/math.py): fp32 graphs use fp32 libm
(`expf`, `tanhf` … through CUDA's overloaded math functions — never fast-math intrinsics), `Maximum`/`Minimum`
propagate NaN, `IntDiv`/`Mod` have Python floor semantics, integer true division goes through double, `Softplus` /
`Log1mexp` use the reference's branch thresholds (scalar/math.py:1253-1282, :1326).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

CTYPE = {
    "bool": "unsigned char", "int8": "signed char", "int16": "short", "int32": "int", "int64": "long long",
    "uint8": "unsigned char", "uint16": "unsigned short", "uint32": "unsigned int", "uint64": "unsigned long long",
    "float32": "float", "float64": "double",
}
ITEMSIZE = {"bool": 1, "int8": 1, "int16": 2, "int32": 4, "int64": 8, "uint8": 1, "uint16": 2, "uint32": 4,
            "uint64": 8, "float32": 4, "float64": 8, "float16": 2}


def is_float(dt: str) -> bool:
    return dt in ("float32", "float64")


def is_int(dt: str) -> bool:
    return dt.startswith("int") or dt.startswith("uint")


def is_uint(dt: str) -> bool:
    return dt.startswith("uint")


class UnsupportedScalarOp(NotImplementedError):
    pass


@dataclass
class ScalarInst:
    op: str  # reference ScalarOp class name, e.g. "Add", "Tanh", "Cast"
    args: list  # refs: ("i", k) input, ("c", k) constant, ("t", k) temp
    in_dtypes: list
    out_dtype: str


@dataclass
class ScalarProgram:
    in_dtypes: list
    out_dtypes: list
    consts: list = field(default_factory=list)  # (dtype, python value)
    insts: list = field(default_factory=list)  # ScalarInst; result of inst k is ("t", k)
    outputs: list = field(default_factory=list)  # refs

    def signature(self) -> str:
        return repr((self.in_dtypes, self.out_dtypes, [(d, repr(v)) for d, v in self.consts],
                     [(i.op, i.args, i.in_dtypes, i.out_dtype) for i in self.insts], self.outputs))

    def n_transcendental(self) -> int:
        heavy = {"Exp", "Exp2", "Expm1", "Log", "Log2", "Log10", "Log1p", "Tanh", "Sinh", "Cosh", "Sin", "Cos", "Tan",
                 "Pow", "Sigmoid", "Softplus", "Log1mexp", "Erf", "Erfc", "Erfcx", "Erfinv", "Erfcinv", "Gamma",
                 "GammaLn", "ArcSin", "ArcCos", "ArcTan", "ArcTan2", "ArcSinh", "ArcCosh", "ArcTanh", "J0", "J1",
                 "I0", "I1", "Psi", "Sqrt", "TrueDiv", "Reciprocal"}
        return sum(1 for i in self.insts if i.op in heavy)


# ---- literals -------------------------------------------------------------------------------------------------------
def literal(dtype: str, value) -> str:
    if dtype == "bool":
        return "((unsigned char)1)" if bool(value) else "((unsigned char)0)"
    if is_int(dtype):
        v = int(value)
        ct = CTYPE[dtype]
        if dtype == "uint64":
            return f"(({ct}){v}ULL)"
        if dtype == "int64":
            if v == -(2 ** 63):
                return "((long long)(-9223372036854775807LL - 1LL))"
            return f"(({ct}){v}LL)"
        return f"(({ct}){v})"
    v = float(value)
    if dtype == "float32":
        if math.isnan(v):
            return "__int_as_float(0x7fc00000)"
        if math.isinf(v):
            return "__int_as_float(0x7f800000)" if v > 0 else "__int_as_float(0xff800000)"
        return f"{np.float32(v).item().hex()}f" if v != 0 else ("-0.0f" if math.copysign(1.0, v) < 0 else "0.0f")
    if dtype == "float64":
        if math.isnan(v):
            return "__longlong_as_double(0x7ff8000000000000LL)"
        if math.isinf(v):
            return "__longlong_as_double(0x7ff0000000000000LL)" if v > 0 else "__longlong_as_double(0xfff0000000000000LL)"
        return f"{v.hex()}" if v != 0 else ("-0.0" if math.copysign(1.0, v) < 0 else "0.0")
    raise UnsupportedScalarOp(f"constant of dtype {dtype}")


# ---- per-op expression table -----------------------------------------------------------------------------------------
def _nan(dt):
    return "__int_as_float(0x7fc00000)" if dt == "float32" else "__longlong_as_double(0x7ff8000000000000LL)"


def _unary_libm(fn):
    def emit(a, it, ot):
        T = CTYPE[ot]
        return f"{fn}(({T})({a[0]}))"
    return emit


def _cmp(sym):
    def emit(a, it, ot):
        return f"((unsigned char)(({a[0]}) {sym} ({a[1]})))"
    return emit


def _add(a, it, ot):
    if ot == "bool":
        return "((unsigned char)(" + " || ".join(f"({x})" for x in a) + "))"
    return "(" + " + ".join(f"({x})" for x in a) + ")"


def _mul(a, it, ot):
    if ot == "bool":
        return "((unsigned char)(" + " && ".join(f"({x})" for x in a) + "))"
    return "(" + " * ".join(f"({x})" for x in a) + ")"


def _truediv(a, it, ot):
    if all(not is_float(t) for t in it):
        return f"(((double)({a[0]})) / ({a[1]}))"
    return f"(({a[0]}) / ({a[1]}))"


def _intdiv(a, it, ot):
    x, y = a
    if is_float(ot):
        return f"floor(({x}) / ({y}))"
    if ot == "bool" or is_uint(ot):
        return f"(({y}) == 0 ? 0 : ({x}) / ({y}))"
    T = CTYPE[ot]
    return (f"ptk_floordiv<{T}>(({T})({x}), ({T})({y}))")


def _mod(a, it, ot):
    x, y = a
    T = CTYPE[ot]
    if is_float(ot):
        return f"ptk_fmod_py<{T}>(({T})({x}), ({T})({y}))"
    if ot == "bool" or is_uint(ot):
        return f"(({y}) == 0 ? 0 : ({x}) % ({y}))"
    return f"ptk_imod_py<{T}>(({T})({x}), ({T})({y}))"


def _maximum(a, it, ot):
    x, y = a
    if ot == "float32" and all(t == "float32" for t in it):
        return f"ptk_max_nan_f32(({x}), ({y}))"  # one FMNMX.NAN: NaN-propagating like the reference's Maximum.c_code
    if is_float(ot):
        return f"((({y}) > ({x})) ? ({y}) : ((({x}) >= ({y})) ? ({x}) : {_nan(ot)}))"
    return f"((({y}) > ({x})) ? ({y}) : ({x}))"


def _minimum(a, it, ot):
    x, y = a
    if ot == "float32" and all(t == "float32" for t in it):
        return f"ptk_min_nan_f32(({x}), ({y}))"
    if is_float(ot):
        return f"((({y}) < ({x})) ? ({y}) : ((({x}) <= ({y})) ? ({x}) : {_nan(ot)}))"
    return f"((({y}) < ({x})) ? ({y}) : ({x}))"


def _abs(a, it, ot):
    if is_float(it[0]):
        return f"fabs({a[0]})"
    if it[0] == "bool" or is_uint(it[0]):
        return f"({a[0]})"
    return f"((({a[0]}) < 0) ? -({a[0]}) : ({a[0]}))"


def _sign(a, it, ot):
    x = a[0]
    if is_float(it[0]):
        T = CTYPE[ot]
        return f"((({x}) > 0) ? ({T})1 : ((({x}) < 0) ? ({T})-1 : (isnan({x}) ? {_nan(ot)} : ({T})0)))"
    if it[0] == "bool" or is_uint(it[0]):
        return f"((({x}) > 0) ? 1 : 0)"
    return f"((({x}) > 0) - (({x}) < 0))"


def _cast(a, it, ot):
    if ot == "bool":
        return f"((unsigned char)(({a[0]}) ? 1 : 0))"
    return f"(({CTYPE[ot]})({a[0]}))"


def _isnan(a, it, ot):
    if is_float(it[0]):
        return f"((unsigned char)(isnan({a[0]}) ? 1 : 0))"
    return "((unsigned char)0)"


def _isinf(a, it, ot):
    if is_float(it[0]):
        return f"((unsigned char)(isinf({a[0]}) ? 1 : 0))"
    return "((unsigned char)0)"


def _invert(a, it, ot):
    if it[0] == "bool":
        return f"((unsigned char)(!({a[0]})))"
    return f"(~({a[0]}))"


def _trunc(a, it, ot):
    return f"trunc({a[0]})"


def _sigmoid(a, it, ot):
    T = CTYPE[ot]
    return f"(({T})1 / (({T})1 + exp(-(({T})({a[0]})))))"


def _softplus(a, it, ot):
    T = CTYPE[ot]
    x = f"(({T})({a[0]}))"
    s = "f" if ot == "float32" else ""
    return (f"(({x} < -37.0{s}) ? exp({x}) : (({x} < 18.0{s}) ? log1p(exp({x})) : "
            f"(({x} < 33.3{s}) ? ({x} + exp(-{x})) : {x})))")


def _log1mexp(a, it, ot):
    T = CTYPE[ot]
    x = f"(({T})({a[0]}))"
    s = "f" if ot == "float32" else ""
    return f"(({x} < -0.6931471805599453{s}) ? log1p(-exp({x})) : log(-expm1({x})))"


def _pow(a, it, ot):
    T = CTYPE[ot]
    if is_float(ot):
        return f"pow(({T})({a[0]}), ({T})({a[1]}))"
    return f"(({T})pow((double)({a[0]}), (double)({a[1]})))"


def _recip(a, it, ot):
    T = CTYPE[ot]
    return f"(({T})1 / ({T})({a[0]}))"


def _round_even(a, it, ot):
    return f"rint({a[0]})"


OPS = {
    "Add": _add, "Mul": _mul,
    "Sub": lambda a, it, ot: f"(({a[0]}) - ({a[1]}))",
    "Neg": lambda a, it, ot: f"(-({a[0]}))",
    "Sqr": lambda a, it, ot: f"(({a[0]}) * ({a[0]}))",
    "Reciprocal": _recip, "TrueDiv": _truediv, "IntDiv": _intdiv, "Mod": _mod, "Pow": _pow,
    "Identity": lambda a, it, ot: f"({a[0]})",
    "Conj": lambda a, it, ot: f"({a[0]})",
    "Second": lambda a, it, ot: f"({a[1]})",
    "Abs": _abs, "Sign": _sign, "Cast": _cast,
    "LT": _cmp("<"), "GT": _cmp(">"), "LE": _cmp("<="), "GE": _cmp(">="), "EQ": _cmp("=="), "NEQ": _cmp("!="),
    "IsNan": _isnan, "IsInf": _isinf,
    "AND": lambda a, it, ot: f"(({a[0]}) & ({a[1]}))",
    "OR": lambda a, it, ot: f"(({a[0]}) | ({a[1]}))",
    "XOR": lambda a, it, ot: f"(({a[0]}) ^ ({a[1]}))",
    "Invert": _invert,
    "Switch": lambda a, it, ot: f"(({a[0]}) ? ({a[1]}) : ({a[2]}))",
    "Clip": lambda a, it, ot: f"((({a[0]}) < ({a[1]})) ? ({a[1]}) : ((({a[0]}) > ({a[2]})) ? ({a[2]}) : ({a[0]})))",
    "Maximum": _maximum, "Minimum": _minimum,
    "Ceil": _unary_libm("ceil"), "Floor": _unary_libm("floor"), "Trunc": _trunc,
    "RoundHalfAwayFromZero": _unary_libm("round"), "RoundHalfToEven": _round_even,
    "Exp": _unary_libm("exp"), "Exp2": _unary_libm("exp2"), "Expm1": _unary_libm("expm1"),
    "Log": _unary_libm("log"), "Log2": _unary_libm("log2"), "Log10": _unary_libm("log10"),
    "Log1p": _unary_libm("log1p"), "Sqrt": _unary_libm("sqrt"),
    "Sin": _unary_libm("sin"), "Cos": _unary_libm("cos"), "Tan": _unary_libm("tan"),
    "ArcSin": _unary_libm("asin"), "ArcCos": _unary_libm("acos"), "ArcTan": _unary_libm("atan"),
    "ArcTan2": lambda a, it, ot: f"atan2(({CTYPE[ot]})({a[0]}), ({CTYPE[ot]})({a[1]}))",
    "Sinh": _unary_libm("sinh"), "Cosh": _unary_libm("cosh"), "Tanh": _unary_libm("tanh"),
    "ArcSinh": _unary_libm("asinh"), "ArcCosh": _unary_libm("acosh"), "ArcTanh": _unary_libm("atanh"),
    "Sigmoid": _sigmoid, "Softplus": _softplus, "Log1mexp": _log1mexp,
    "Erf": _unary_libm("erf"), "Erfc": _unary_libm("erfc"), "Erfcx": _unary_libm("erfcx"),
    "Erfinv": _unary_libm("erfinv"), "Erfcinv": _unary_libm("erfcinv"),
    "Gamma": _unary_libm("tgamma"), "GammaLn": _unary_libm("lgamma"),
    "J0": _unary_libm("j0"), "J1": _unary_libm("j1"),
    "I0": _unary_libm("cyl_bessel_i0"), "I1": _unary_libm("cyl_bessel_i1"),
}

PRELUDE = r"""
// ---- ptk scalar helpers ----
__device__ __forceinline__ float ptk_max_nan_f32(float a, float b) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ptk_min_nan_f32(float a, float b) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
// Python floor-division / modulo semantics of IntDiv / Mod
template <typename T> __device__ __forceinline__ T ptk_floordiv(T x, T y) {
  if (y == 0) return 0;
  T q = x / y;
  if ((x % y != 0) && ((x < 0) != (y < 0))) --q;
  return q;
}
template <typename T> __device__ __forceinline__ T ptk_imod_py(T x, T y) {
  if (y == 0) return 0;
  T r = x % y;
  if (r != 0 && ((r < 0) != (y < 0))) r += y;
  return r;
}
template <typename T> __device__ __forceinline__ T ptk_fmod_py(T x, T y) {
  if (y == 0) return x - x + (T)__int_as_float(0x7fc00000);
  T r = fmod(x, y);
  if (r != 0 && ((r < 0) != (y < 0))) r += y;
  return r;
}
"""


def _ref_expr(ref, prog: ScalarProgram) -> str:
    kind, k = ref
    if kind == "i":
        return f"i{k}"
    if kind == "c":
        d, v = prog.consts[k]
        return literal(d, v)
    return f"t{k}"


def emit_body(prog: ScalarProgram, fn_name: str = "ptk_body") -> str:
    """`__device__ void fn(const T0& i0, ..., O0& o0, ...)` computing all outputs of `prog`."""
    for d in list(prog.in_dtypes) + list(prog.out_dtypes) + [i.out_dtype for i in prog.insts]:
        if d not in CTYPE:
            raise UnsupportedScalarOp(f"dtype {d} has no device path (the reference's C linker has none for float16/complex either)")
    params = [f"const {CTYPE[d]} i{k}" for k, d in enumerate(prog.in_dtypes)]
    params += [f"{CTYPE[d]}& o{k}" for k, d in enumerate(prog.out_dtypes)]
    lines = [f"__device__ __forceinline__ void {fn_name}({', '.join(params)}) {{"]
    for k, inst in enumerate(prog.insts):
        fn = OPS.get(inst.op)
        if fn is None:
            raise UnsupportedScalarOp(f"scalar op {inst.op} has no sm_100a device expression yet")
        args = [_ref_expr(r, prog) for r in inst.args]
        expr = fn(args, inst.in_dtypes, inst.out_dtype)
        T = CTYPE[inst.out_dtype]
        lines.append(f"  const {T} t{k} = ({T})({expr});")
    for k, ref in enumerate(prog.outputs):
        lines.append(f"  o{k} = ({CTYPE[prog.out_dtypes[k]]})({_ref_expr(ref, prog)});")
    lines.append("}")
    return "\n".join(lines)


def single_op_program(op: str, in_dtypes, out_dtype) -> ScalarProgram:
    return ScalarProgram(
        in_dtypes=list(in_dtypes), out_dtypes=[out_dtype],
        insts=[ScalarInst(op, [("i", k) for k in range(len(in_dtypes))], list(in_dtypes), out_dtype)],
        outputs=[("t", 0)],
    )
