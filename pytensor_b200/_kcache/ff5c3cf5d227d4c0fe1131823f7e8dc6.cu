
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
__device__ __forceinline__ ACC ptk_red(ACC a, ACC b) { return (ACC)((a) + (b)); }
extern "C" __global__ void __launch_bounds__(256) ptk_red_finish_ac870bbe928d3456(const ACC* __restrict__ part, OUT* __restrict__ out,
                                                         long long n_out, int nsplit, long long part_stride_o,
                                                         long long part_stride_s) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long o = warp; o < n_out; o += nwarps) {
    ACC acc = (ACC)0.0;
    for (int s = lane; s < nsplit; s += 32) acc = ptk_red(acc, part[o * part_stride_o + s * part_stride_s]);
    #pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc = ptk_red(acc, ptk_shfl_xor<ACC>(acc, m));
    if (lane == 0) out[o] = (OUT)acc;
  }
}
