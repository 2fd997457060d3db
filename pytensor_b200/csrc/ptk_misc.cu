// Glue kernels of the "more of the Op library" row: ARange, Argmax, CumOp.
// All HBM-bound integer/byte style work: coalesced access along the contiguous axis, one pass over the input.
#include <algorithm>
#include <limits>
#include <type_traits>

#include "ptk_common.h"

namespace {

using ptk::fail;

inline unsigned grid_for(int64_t work_items, int per_block) {
  int64_t want = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)std::max(1, ptk::sm_count()) * 16;
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(want, cap));
}

// ---- ARange: out[i] = first + i * delta, evaluated in the output type like NumPy's <type>_fill ------------------------
template <typename T>
__global__ void arange_float_kernel(T* __restrict__ out, int64_t n, T first, T delta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if constexpr (sizeof(T) == 4) {
      out[i] = __fadd_rn(first, __fmul_rn((float)i, delta));  // no FMA contraction: two roundings, as the C loop does
    } else {
      out[i] = __dadd_rn(first, __dmul_rn((double)i, delta));
    }
  }
}
template <typename T>
__global__ void arange_int_kernel(T* __restrict__ out, int64_t n, int64_t first, int64_t delta) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (T)((uint64_t)first + (uint64_t)i * (uint64_t)delta);  // wraps like the fixed-width C arithmetic
}

// ---- Argmax --------------------------------------------------------------------------------------------------------
// np.argmax semantics: index of the FIRST maximal element; a NaN counts as maximal (the first NaN wins).
template <typename T>
__device__ __forceinline__ bool is_nan_v(T v) {
  if constexpr (std::is_floating_point<T>::value) return v != v;
  return false;
}
template <typename T>
__device__ __forceinline__ bool better(T v, int64_t i, T bv, int64_t bi) {
  const bool n1 = is_nan_v(v), n2 = is_nan_v(bv);
  if (n1 || n2) return n1 && (!n2 || i < bi);
  return v > bv || (v == bv && i < bi);
}

// inner == 1: one warp (small rows) or one CTA (long rows / few rows) per row; lanes stride over the row (coalesced)
template <typename T, int THREADS>
__global__ void argmax_rows_kernel(const T* __restrict__ x, int64_t* __restrict__ out, int64_t rows, int64_t n) {
  constexpr int WARPS = THREADS / 32;
  __shared__ T s_v[WARPS];
  __shared__ int64_t s_i[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* row = x + r * n;
    T bv = row[0];
    int64_t bi = 0;
    for (int64_t j = threadIdx.x; j < n; j += THREADS) {
      const T v = row[j];
      if (better(v, j, bv, bi)) { bv = v; bi = j; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const T ov = __shfl_down_sync(0xffffffffu, bv, off);
      const int64_t oi = __shfl_down_sync(0xffffffffu, bi, off);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (WARPS > 1) {
      if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
      __syncthreads();
      if (warp == 0) {
        bv = s_v[lane < WARPS ? lane : 0];
        bi = s_i[lane < WARPS ? lane : 0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const T ov = __shfl_down_sync(0xffffffffu, bv, off);
          const int64_t oi = __shfl_down_sync(0xffffffffu, bi, off);
          if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
      }
    }
    if (threadIdx.x == 0) out[r] = bi;
    if (WARPS > 1) __syncthreads();
  }
}

// inner > 1: one thread per output (o, i); consecutive threads read consecutive addresses at every step j
template <typename T>
__global__ void argmax_cols_kernel(const T* __restrict__ x, int64_t* __restrict__ out, int64_t outer, int64_t n,
                                   int64_t inner) {
  const int64_t total = outer * inner;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = t / inner, i = t - o * inner;
    const T* p = x + o * n * inner + i;
    T bv = p[0];
    int64_t bi = 0;
    for (int64_t j = 1; j < n; ++j) {
      const T v = p[j * inner];
      if (better(v, j, bv, bi)) { bv = v; bi = j; }
    }
    out[t] = bi;
  }
}

template <typename T>
ptk_status argmax_t(const void* x, int64_t* out, int64_t outer, int64_t n, int64_t inner, cudaStream_t st) {
  if (inner == 1) {
    if (n >= 2048 || outer < 64) {
      const unsigned g = (unsigned)std::min<int64_t>(outer, (int64_t)std::max(1, ptk::sm_count()) * 8);
      argmax_rows_kernel<T, 256><<<g, 256, 0, st>>>((const T*)x, out, outer, n);
    } else {
      const unsigned g = (unsigned)std::min<int64_t>(outer, (int64_t)std::max(1, ptk::sm_count()) * 64);
      argmax_rows_kernel<T, 32><<<g, 32, 0, st>>>((const T*)x, out, outer, n);
    }
  } else {
    argmax_cols_kernel<T><<<grid_for(outer * inner, 256), 256, 0, st>>>((const T*)x, out, outer, n, inner);
  }
  PTK_LAUNCH_CHECK("argmax");
  return PTK_OK;
}

// ---- CumOp ---------------------------------------------------------------------------------------------------------
template <typename T, int OP>
__device__ __forceinline__ T cum_combine(T a, T b) {
  return OP == 0 ? (T)(a + b) : (T)(a * b);
}

// inner > 1: one thread per line, strictly sequential along the axis (same order as np.add.accumulate -> bit-exact)
template <typename T, int OP>
__global__ void cum_cols_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t outer, int64_t n, int64_t inner) {
  const int64_t total = outer * inner;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = t / inner, i = t - o * inner;
    const T* p = x + o * n * inner + i;
    T* q = out + o * n * inner + i;
    T acc = p[0];
    q[0] = acc;
    for (int64_t j = 1; j < n; ++j) {
      acc = cum_combine<T, OP>(acc, p[j * inner]);
      q[j * inner] = acc;
    }
  }
}

// inner == 1: one warp per row; 32-element chunks scanned with shuffles, the running total carried between chunks
// (tree order inside a chunk: equal to the sequential result for integers, within rounding for floats)
template <typename T, int OP>
__global__ void cum_rows_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t rows, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const T ident = OP == 0 ? (T)0 : (T)1;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const T* row = x + r * n;
    T* orow = out + r * n;
    T carry = ident;
    bool first = true;
    for (int64_t base = 0; base < n; base += 32) {
      const int64_t j = base + lane;
      T v = j < n ? row[j] : ident;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const T u = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v = cum_combine<T, OP>(u, v);
      }
      if (!first) v = cum_combine<T, OP>(carry, v);
      if (j < n) orow[j] = v;
      carry = __shfl_sync(0xffffffffu, v, 31);
      first = false;
    }
  }
}

template <typename T>
ptk_status cumop_t(int op, const void* x, void* out, int64_t outer, int64_t n, int64_t inner, cudaStream_t st) {
  if (inner == 1) {
    const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((outer + 7) / 8, (int64_t)std::max(1, ptk::sm_count()) * 16));
    if (op == 0) cum_rows_kernel<T, 0><<<g, 256, 0, st>>>((const T*)x, (T*)out, outer, n);
    else cum_rows_kernel<T, 1><<<g, 256, 0, st>>>((const T*)x, (T*)out, outer, n);
  } else {
    const unsigned g = grid_for(outer * inner, 256);
    if (op == 0) cum_cols_kernel<T, 0><<<g, 256, 0, st>>>((const T*)x, (T*)out, outer, n, inner);
    else cum_cols_kernel<T, 1><<<g, 256, 0, st>>>((const T*)x, (T*)out, outer, n, inner);
  }
  PTK_LAUNCH_CHECK("cumop");
  return PTK_OK;
}

// ---- linear index of several integer index arrays (advanced indexing on consecutive axes) -----------------------------
struct LinIdxArgs {
  const int64_t* idx[8];
  int64_t dim[8];
  int64_t stride[8];
  int k;
};
__global__ void linearize_index_kernel(LinIdxArgs a, int64_t n, int64_t* __restrict__ out, int* __restrict__ err) {
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t lin = 0;
    bool bad = false;
    for (int j = 0; j < a.k; ++j) {
      int64_t v = a.idx[j][t];
      if (v < 0) v += a.dim[j];            // NumPy's negative-index wrap, per axis
      if (v < 0 || v >= a.dim[j]) { bad = true; v = 0; }
      lin += v * a.stride[j];
    }
    if (bad) *err = 1;
    out[t] = lin;
  }
}


// ---- Nonzero (boolean-mask indexing): ascending flat positions of the set bytes of a mask -------------------------------
// Three passes over NZ_TILE-byte tiles: per-tile counts (16-byte loads, popcount of the "byte != 0" bits), an exclusive scan
// of the tile counts by one CTA (total appended), and an ordered per-tile compaction (warp ballots keep positions ascending).
constexpr int NZ_THREADS = 256;
constexpr int NZ_TILE = NZ_THREADS * 16;  // bytes of mask per tile

__device__ __forceinline__ int nz_count16(const uint8_t* __restrict__ m, int64_t base, int64_t n) {
  int c = 0;
  if (base + 16 <= n && ((reinterpret_cast<uintptr_t>(m + base) & 15) == 0)) {
    const uint4 v = *reinterpret_cast<const uint4*>(m + base);
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // a byte is "set" when any of its bits is: fold each byte's bits into its lowest bit, then count
      unsigned t = w[j];
      t |= t >> 4; t |= t >> 2; t |= t >> 1;
      c += __popc(t & 0x01010101u);
    }
  } else {
    for (int j = 0; j < 16; ++j)
      if (base + j < n && m[base + j]) ++c;
  }
  return c;
}

__global__ void __launch_bounds__(NZ_THREADS) nonzero_count_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                    int64_t* __restrict__ tile_count, int64_t tiles) {
  __shared__ int s_part[NZ_THREADS / 32];
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    int c = nz_count16(mask, tile * NZ_TILE + (int64_t)threadIdx.x * 16, n);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
#pragma unroll
      for (int w = 0; w < NZ_THREADS / 32; ++w) t += s_part[w];
      tile_count[tile] = t;
    }
    __syncthreads();
  }
}

// in place: tile_count[i] <- sum of tile_count[0..i) ; tile_count[tiles] <- total
__global__ void __launch_bounds__(1024) nonzero_scan_kernel(int64_t* __restrict__ tile_count, int64_t tiles) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = 0; base < tiles; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < tiles ? tile_count[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane], winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t u = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += u;
      }
      s_warp[lane] = winc - w;  // exclusive offset of each warp
    }
    __syncthreads();
    const int64_t carry = s_carry;
    if (i < tiles) tile_count[i] = carry + s_warp[warp] + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) tile_count[tiles] = s_carry;
}

__global__ void __launch_bounds__(NZ_THREADS) nonzero_fill_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                                   const int64_t* __restrict__ tile_offset, int64_t tiles,
                                                                   int64_t* __restrict__ out) {
  __shared__ int s_part[NZ_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    // thread t owns bytes [16 t, 16 t + 16) of the tile: count, exclusive scan over the CTA, then write in order
    const int64_t base = tile * NZ_TILE + (int64_t)threadIdx.x * 16;
    const int c = nz_count16(mask, base, n);
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < NZ_THREADS / 32; ++w)
      if (w < warp) before += s_part[w];
    int64_t pos = tile_offset[tile] + before + inc - c;
    if (c) {
      for (int j = 0; j < 16; ++j)
        if (base + j < n && mask[base + j]) out[pos++] = base + j;
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" ptk_status ptk_linearize_index(int k, const void* const* idx, const int64_t* dims, int64_t n, int64_t* out,
                                          int* err_flag, void* stream) {
  PTK_REQUIRE_INIT();
  if (k < 1 || k > 8) return fail(PTK_ERR_ARG, "ptk_linearize_index: 1..8 index arrays");
  if (n <= 0) return PTK_OK;
  LinIdxArgs a;
  a.k = k;
  int64_t stride = 1;
  for (int j = k - 1; j >= 0; --j) {
    a.idx[j] = (const int64_t*)idx[j];
    a.dim[j] = dims[j];
    a.stride[j] = stride;
    stride *= dims[j];
  }
  linearize_index_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, n, out, err_flag);
  PTK_LAUNCH_CHECK("linearize_index");
  return PTK_OK;
}

extern "C" ptk_status ptk_arange(int dtype, void* out, int64_t n, double first_f, double delta_f, int64_t first_i,
                                 int64_t delta_i, void* stream) {
  PTK_REQUIRE_INIT();
  if (n <= 0) return PTK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = grid_for(n, 256 * 4);
  switch (dtype) {
    case PTK_F32: arange_float_kernel<float><<<g, 256, 0, st>>>((float*)out, n, (float)first_f, (float)delta_f); break;
    case PTK_F64: arange_float_kernel<double><<<g, 256, 0, st>>>((double*)out, n, first_f, delta_f); break;
    case PTK_I8: arange_int_kernel<int8_t><<<g, 256, 0, st>>>((int8_t*)out, n, first_i, delta_i); break;
    case PTK_U8: arange_int_kernel<uint8_t><<<g, 256, 0, st>>>((uint8_t*)out, n, first_i, delta_i); break;
    case PTK_I16: arange_int_kernel<int16_t><<<g, 256, 0, st>>>((int16_t*)out, n, first_i, delta_i); break;
    case PTK_U16: arange_int_kernel<uint16_t><<<g, 256, 0, st>>>((uint16_t*)out, n, first_i, delta_i); break;
    case PTK_I32: arange_int_kernel<int32_t><<<g, 256, 0, st>>>((int32_t*)out, n, first_i, delta_i); break;
    case PTK_U32: arange_int_kernel<uint32_t><<<g, 256, 0, st>>>((uint32_t*)out, n, first_i, delta_i); break;
    case PTK_I64: arange_int_kernel<int64_t><<<g, 256, 0, st>>>((int64_t*)out, n, first_i, delta_i); break;
    case PTK_U64: arange_int_kernel<uint64_t><<<g, 256, 0, st>>>((uint64_t*)out, n, first_i, delta_i); break;
    default: return fail(PTK_ERR_ARG, "ptk_arange: unsupported dtype");
  }
  PTK_LAUNCH_CHECK("arange");
  return PTK_OK;
}

extern "C" ptk_status ptk_argmax(int dtype, const void* x, int64_t* out, int64_t outer, int64_t n, int64_t inner,
                                 void* stream) {
  PTK_REQUIRE_INIT();
  if (outer * inner == 0) return PTK_OK;
  if (n <= 0) return fail(PTK_ERR_ARG, "ptk_argmax: attempt to get argmax of an empty sequence");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case PTK_BOOL: case PTK_U8: return argmax_t<uint8_t>(x, out, outer, n, inner, st);
    case PTK_I8: return argmax_t<int8_t>(x, out, outer, n, inner, st);
    case PTK_I16: return argmax_t<int16_t>(x, out, outer, n, inner, st);
    case PTK_U16: return argmax_t<uint16_t>(x, out, outer, n, inner, st);
    case PTK_I32: return argmax_t<int32_t>(x, out, outer, n, inner, st);
    case PTK_U32: return argmax_t<uint32_t>(x, out, outer, n, inner, st);
    case PTK_I64: return argmax_t<int64_t>(x, out, outer, n, inner, st);
    case PTK_U64: return argmax_t<uint64_t>(x, out, outer, n, inner, st);
    case PTK_F32: return argmax_t<float>(x, out, outer, n, inner, st);
    case PTK_F64: return argmax_t<double>(x, out, outer, n, inner, st);
  }
  return fail(PTK_ERR_ARG, "ptk_argmax: unsupported dtype");
}

extern "C" ptk_status ptk_cumop(int dtype, int op, const void* x, void* out, int64_t outer, int64_t n, int64_t inner,
                                void* stream) {
  PTK_REQUIRE_INIT();
  if (outer * n * inner == 0) return PTK_OK;
  if (op != 0 && op != 1) return fail(PTK_ERR_ARG, "ptk_cumop: op must be 0 (add) or 1 (mul)");
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case PTK_F32: return cumop_t<float>(op, x, out, outer, n, inner, st);
    case PTK_F64: return cumop_t<double>(op, x, out, outer, n, inner, st);
    case PTK_I64: return cumop_t<int64_t>(op, x, out, outer, n, inner, st);
    case PTK_U64: return cumop_t<uint64_t>(op, x, out, outer, n, inner, st);
  }
  return fail(PTK_ERR_ARG, "ptk_cumop: dtype must be float32/float64/int64/uint64 (np.cumsum keeps only those)");
}

extern "C" size_t ptk_nonzero_workspace_bytes(int64_t n) {
  const int64_t tiles = (n + NZ_TILE - 1) / NZ_TILE;
  return (size_t)(tiles + 1) * sizeof(int64_t);
}

extern "C" ptk_status ptk_nonzero_count(const void* mask, int64_t n, void* workspace, size_t workspace_bytes, void* stream) {
  PTK_REQUIRE_INIT();
  if (n < 0) return fail(PTK_ERR_ARG, "ptk_nonzero_count: negative length");
  if (workspace == nullptr || workspace_bytes < ptk_nonzero_workspace_bytes(n))
    return fail(PTK_ERR_ARG, "ptk_nonzero_count: workspace too small (see ptk_nonzero_workspace_bytes)");
  const int64_t tiles = (n + NZ_TILE - 1) / NZ_TILE;
  cudaStream_t st = (cudaStream_t)stream;
  if (tiles) {
    nonzero_count_kernel<<<grid_for(tiles, 1), NZ_THREADS, 0, st>>>((const uint8_t*)mask, n, (int64_t*)workspace, tiles);
    PTK_LAUNCH_CHECK("nonzero_count");
  }
  nonzero_scan_kernel<<<1, 1024, 0, st>>>((int64_t*)workspace, tiles);
  PTK_LAUNCH_CHECK("nonzero_scan");
  return PTK_OK;
}

extern "C" ptk_status ptk_nonzero_fill(const void* mask, int64_t n, const void* workspace, int64_t* out, void* stream) {
  PTK_REQUIRE_INIT();
  const int64_t tiles = (n + NZ_TILE - 1) / NZ_TILE;
  if (tiles <= 0) return PTK_OK;
  nonzero_fill_kernel<<<grid_for(tiles, 1), NZ_THREADS, 0, (cudaStream_t)stream>>>((const uint8_t*)mask, n,
                                                                                    (const int64_t*)workspace, tiles, out);
  PTK_LAUNCH_CHECK("nonzero_fill");
  return PTK_OK;
}
