
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

__device__ __forceinline__ void ptk_body(const float i0, const long long i1, const long long i2, float& o0, float& o1) {
  const unsigned char t0 = (unsigned char)(((unsigned char)((i1) >= (i2))));
  const float t1 = (float)(((float)(t0)));
  const float t2 = (float)(((i0) * (t1)));
  const float t3 = (float)(((0x1.0000000000000p+0f) - (t1)));
  const float t4 = (float)(((i0) * (t3)));
  o0 = (float)(t4);
  o1 = (float)(t2);
}

struct EwDims { int ndim; long long shape[8]; long long st[5][8]; };

extern "C" __global__ void __launch_bounds__(256) ptk_ew_gen_384a9b3cd9a69505(const float* pi0, const long long* pi1, const long long* pi2, float* po0, float* po1, const EwDims dims, long long total) {
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    long long rem = i;
    long long off0 = 0;
    long long off1 = 0;
    long long off2 = 0;
    long long off3 = 0;
    long long off4 = 0;
#pragma unroll
    for (int k = 8 - 1; k >= 0; --k) {
      if (k < dims.ndim) {
        const long long q = rem / dims.shape[k];
        const long long cidx = rem - q * dims.shape[k];
        rem = q;
        off0 += cidx * dims.st[0][k];
        off1 += cidx * dims.st[1][k];
        off2 += cidx * dims.st[2][k];
        off3 += cidx * dims.st[3][k];
        off4 += cidx * dims.st[4][k];
      }
    }
    float to0;
    float to1;
    ptk_body(pi0[off0], pi1[off1], pi2[off2], to0, to1);
    po0[off3] = to0;
    po1[off4] = to1;
  }
}
