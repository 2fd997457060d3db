
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


template <typename T> __device__ __forceinline__ T ptk_nanmax(T a, T b) { return (b > a) ? b : ((a >= b) ? a : (a + b)); }
template <typename T> __device__ __forceinline__ T ptk_nanmin(T a, T b) { return (b < a) ? b : ((a <= b) ? a : (a + b)); }
template <typename T> __device__ __forceinline__ T ptk_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ unsigned char ptk_shfl_xor<unsigned char>(unsigned char v, int m) { return (unsigned char)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ signed char ptk_shfl_xor<signed char>(signed char v, int m) { return (signed char)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ short ptk_shfl_xor<short>(short v, int m) { return (short)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ unsigned short ptk_shfl_xor<unsigned short>(unsigned short v, int m) { return (unsigned short)__shfl_xor_sync(0xffffffffu, (int)v, m); }

typedef double ACC;
typedef double OUT;
__device__ __forceinline__ ACC ptk_red(ACC a, ACC b) { return (ACC)((a) + (b)); }
struct RdDims { int nk; int nr; long long kshape[8]; long long kst[8]; long long rshape[8]; long long rst[8]; };
extern "C" __global__ void __launch_bounds__(256) ptk_red_generic_8887906101501940(const double* __restrict__ in, OUT* __restrict__ out,
                                                         const RdDims d, long long n_out, long long n_red) {
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += gstride) {
    long long rem = o, base = 0;
    for (int k = d.nk - 1; k >= 0; --k) { const long long q = rem / d.kshape[k]; base += (rem - q * d.kshape[k]) * d.kst[k]; rem = q; }
    ACC acc = (ACC)0.0;
    for (long long j = 0; j < n_red; ++j) {
      long long rj = j, off = base;
      for (int k = d.nr - 1; k >= 0; --k) { const long long q = rj / d.rshape[k]; off += (rj - q * d.rshape[k]) * d.rst[k]; rj = q; }
      acc = ptk_red(acc, (ACC)in[off]);
    }
    out[o] = (OUT)acc;
  }
}
