"""CUDA source generation for fused Elemwise kernels (K1) — one kernel per `Elemwise(Composite)` node.

Replaces what the C linker emits in `Elemwise._c_all` (pytensor/tensor/elemwise.py:848-1167: contiguous fast path
:1118-1146, otherwise the stride-sorted loop nests of elemwise_cgen.py:294-465).  Two kernel shapes:

* `vec`  — the HBM-bound shape.  After the host collapsed the iteration space to (rows, cols), every operand is
  either contiguous along `cols` (128-bit `ld.global.v4` / `st.global.v4` per 4-byte element type, U chunks in
  flight per thread) or constant along `cols` (row/scalar broadcast: one scalar load per chunk).  Grid-stride,
  sized by the host to a multiple of the SM count.
* `gen`  — any <= 8-d strided/broadcast pattern, one element per thread-iteration (index arithmetic per element).

The scalar body is `codegen.scalar.emit_body(prog)`.
"""

from __future__ import annotations

from .scalar import CTYPE, ITEMSIZE, PRELUDE, ScalarProgram, emit_body

MAX_DIMS = 8
VEC_UNROLL = 4

_VEC_HELPERS = r"""
template <typename T, int N> struct __align__(sizeof(T) * N) PVec { T v[N]; };
template <typename T, int N> __device__ __forceinline__ PVec<T, N> ptk_ldv(const T* p) {
  return *reinterpret_cast<const PVec<T, N>*>(p);
}
template <typename T, int N> __device__ __forceinline__ void ptk_stv(T* p, const PVec<T, N>& v) {
  *reinterpret_cast<PVec<T, N>*>(p) = v;
}
__device__ __forceinline__ void ptk_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// A load the compiler may not move: issued exactly where it is written (`asm volatile`), so that the software pipeline of
// the fused map+row-reduce kernel really has the NEXT trip's data in flight while the current trip is computed (left to
// itself the compiler sinks plain loads down to their first use to save registers).
template <typename T, int N> __device__ __forceinline__ PVec<T, N> ptk_ldv_pin(const T* p) {
  PVec<T, N> r;
  if constexpr (sizeof(PVec<T, N>) % 16 == 0) {
    uint4* d = reinterpret_cast<uint4*>(&r);
    const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PVec<T, N>) / 16); ++i)
      asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(d[i].x), "=r"(d[i].y), "=r"(d[i].z), "=r"(d[i].w) : "l"(s + i));
  } else {
    r = *reinterpret_cast<const PVec<T, N>*>(p);
  }
  return r;
}
"""


def vec_width(dtypes) -> int:
    """Elements per thread-chunk: 16 bytes of the widest... narrowest useful: 4 for 4/8-byte types, 8/16 for small."""
    m = max(ITEMSIZE[d] for d in dtypes)
    if m >= 4:
        return 4
    if m == 2:
        return 8
    return 16


def gen_vec_kernel(prog: ScalarProgram, name: str, col_modes: tuple, inplace: dict, vw: int,
                   unroll: int = VEC_UNROLL, flat: bool = False) -> str:
    """col_modes[k] in {0,1} for the n_in inputs followed by the n_out outputs (outputs are always 1).
    inplace: {out_idx: in_idx}.  Kernel params: pointers..., row strides..., nchunks, chunks_per_row, n_tail_start,
    n_total (flat tail handling: elements [n_tail_start, n_total) are done one by one when rows == 1)."""
    n_in, n_out = len(prog.in_dtypes), len(prog.out_dtypes)
    restrict = "" if inplace else " __restrict__"
    params = []
    for k, d in enumerate(prog.in_dtypes):
        params.append(f"const {CTYPE[d]}*{restrict} pi{k}")
    for k, d in enumerate(prog.out_dtypes):
        params.append(f"{CTYPE[d]}*{restrict} po{k}")
    params += [f"long long rsi{k}" for k in range(n_in)]
    params += [f"long long rso{k}" for k in range(n_out)]
    params += ["long long nchunks", "unsigned int cpr", "long long tail_start", "long long n_total"]

    loads, scal = [], []
    for k, d in enumerate(prog.in_dtypes):
        T = CTYPE[d]
        if col_modes[k] == 1:
            loads.append(f"      PVec<{T}, VW> vi{k}[U];")
        else:
            loads.append(f"      {T} vi{k}[U];")
    ld_body = []
    for k, d in enumerate(prog.in_dtypes):
        T = CTYPE[d]
        if flat:
            if col_modes[k] == 1:
                ld_body.append(f"          vi{k}[u] = ptk_ldv<{T}, VW>(pi{k} + q * VW);")
            else:
                ld_body.append(f"          vi{k}[u] = pi{k}[0];")
        elif col_modes[k] == 1:
            ld_body.append(f"          vi{k}[u] = ptk_ldv<{T}, VW>(pi{k} + r * rsi{k} + c);")
        else:
            ld_body.append(f"          vi{k}[u] = pi{k}[r * rsi{k}];")
    call_in = []
    for k in range(n_in):
        call_in.append(f"vi{k}[u].v[e]" if col_modes[k] == 1 else f"vi{k}[u]")
    out_decl = "\n".join(f"          PVec<{CTYPE[d]}, VW> vo{k};" for k, d in enumerate(prog.out_dtypes))
    call_out = [f"vo{k}.v[e]" for k in range(n_out)]
    if flat:
        st_body = "\n".join(f"          ptk_stv<{CTYPE[d]}, VW>(po{k} + q * VW, vo{k});"
                            for k, d in enumerate(prog.out_dtypes))
    else:
        st_body = "\n".join(f"          ptk_stv<{CTYPE[d]}, VW>(po{k} + r * rso{k} + c, vo{k});"
                            for k, d in enumerate(prog.out_dtypes))
    tail_in = [f"pi{k}[i]" if col_modes[k] == 1 else f"pi{k}[0]" for k in range(n_in)]
    tail_tmp = "\n".join(f"      {CTYPE[d]} to{k};" for k, d in enumerate(prog.out_dtypes))
    tail_st = "\n".join(f"      po{k}[i] = to{k};" for k in range(n_out))

    ROWCOL = """          long long r, c;
          if (nchunks < 0x7fffffffLL) { unsigned int qq = (unsigned int)q; unsigned int r32 = qq / cpr; r = r32; c = (long long)(qq - r32 * cpr) * VW; }
          else { r = q / cpr; c = (q - r * cpr) * VW; }
          rr[u] = r; cc[u] = c;"""
    return f"""{PRELUDE}
{_VEC_HELPERS}
{emit_body(prog)}

#define VW {vw}
#define U {unroll}
extern "C" __global__ void __launch_bounds__(256) {name}({', '.join(params)}) {{
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long base = gtid; base < nchunks; base += gstride * U) {{
{chr(10).join(loads)}
{'' if flat else '      long long rr[U]; long long cc[U];'}
#pragma unroll
      for (int u = 0; u < U; ++u) {{
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {{
{'' if flat else ROWCOL}
{chr(10).join(ld_body)}
        }}
      }}
#pragma unroll
      for (int u = 0; u < U; ++u) {{
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {{
{'' if flat else '          const long long r = rr[u], c = cc[u];'}
{out_decl}
#pragma unroll
          for (int e = 0; e < VW; ++e) {{
            ptk_body({', '.join(call_in + call_out)});
          }}
{st_body}
        }}
      }}
  }}
  // flat tail (rows == 1 only): the last n_total % VW elements
  for (long long i = tail_start + gtid; i < n_total; i += gstride) {{
{tail_tmp}
      ptk_body({', '.join(tail_in + [f'to{k}' for k in range(n_out)])});
{tail_st}
  }}
}}
"""


def gen_generic_kernel(prog: ScalarProgram, name: str, inplace: dict) -> str:
    """Any-stride kernel. Params: pointers..., EwDims (by value), total."""
    n_in, n_out = len(prog.in_dtypes), len(prog.out_dtypes)
    nops = n_in + n_out
    restrict = "" if inplace else " __restrict__"
    params = []
    for k, d in enumerate(prog.in_dtypes):
        params.append(f"const {CTYPE[d]}*{restrict} pi{k}")
    for k, d in enumerate(prog.out_dtypes):
        params.append(f"{CTYPE[d]}*{restrict} po{k}")
    params += ["const EwDims dims", "long long total"]
    offs_decl = "\n".join(f"    long long off{j} = 0;" for j in range(nops))
    offs_acc = "\n".join(f"        off{j} += cidx * dims.st[{j}][k];" for j in range(nops))
    call_in = [f"pi{k}[off{k}]" for k in range(n_in)]
    tmp_decl = "\n".join(f"    {CTYPE[d]} to{k};" for k, d in enumerate(prog.out_dtypes))
    stores = "\n".join(f"    po{k}[off{n_in + k}] = to{k};" for k in range(n_out))
    return f"""{PRELUDE}
{emit_body(prog)}

struct EwDims {{ int ndim; long long shape[{MAX_DIMS}]; long long st[{nops}][{MAX_DIMS}]; }};

extern "C" __global__ void __launch_bounds__(256) {name}({', '.join(params)}) {{
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {{
    long long rem = i;
{offs_decl}
#pragma unroll
    for (int k = {MAX_DIMS} - 1; k >= 0; --k) {{
      if (k < dims.ndim) {{
        long long cidx = rem;                 // the outermost axis (k == 0) keeps what is left: no division
        if (k > 0) {{
          const long long q = rem / dims.shape[k];
          cidx = rem - q * dims.shape[k];
          rem = q;
        }}
{offs_acc}
      }}
    }}
{tmp_decl}
    ptk_body({', '.join(call_in + [f'to{k}' for k in range(n_out)])});
{stores}
  }}
}}
"""
