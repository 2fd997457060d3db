
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
typedef float OUT;
__device__ __forceinline__ ACC ptk_red(ACC a, ACC b) { return (ACC)ptk_nanmax((a), (b)); }
extern "C" __global__ void __launch_bounds__(256) ptk_red_col_f6455841884ab4d7(const float* __restrict__ in, void* __restrict__ outp,
                                                         long long outer, long long red, long long inner, int nsplit) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= inner) return;
  const long long per = (red + nsplit - 1) / nsplit;
  const long long r_lo = (long long)blockIdx.z * per;
  const long long r_hi = (r_lo + per < red) ? (r_lo + per) : red;
  for (long long o = blockIdx.y; o < outer; o += gridDim.y) {
    const float* p = in + o * red * inner + i;
    ACC a0 = (ACC)__longlong_as_double(0xfff0000000000000LL), a1 = a0, a2 = a0, a3 = a0;
    long long r = r_lo;
    for (; r + 3 < r_hi; r += 4) {
      const float x0 = p[r * inner], x1 = p[(r + 1) * inner], x2 = p[(r + 2) * inner], x3 = p[(r + 3) * inner];
      a0 = ptk_red(a0, (ACC)x0); a1 = ptk_red(a1, (ACC)x1); a2 = ptk_red(a2, (ACC)x2); a3 = ptk_red(a3, (ACC)x3);
    }
    for (; r < r_hi; ++r) a0 = ptk_red(a0, (ACC)p[r * inner]);
    const ACC acc = ptk_red(ptk_red(a0, a1), ptk_red(a2, a3));
    if (nsplit == 1) reinterpret_cast<OUT*>(outp)[o * inner + i] = (OUT)acc;
    else reinterpret_cast<ACC*>(outp)[((long long)blockIdx.z * outer + o) * inner + i] = acc;
  }
}
