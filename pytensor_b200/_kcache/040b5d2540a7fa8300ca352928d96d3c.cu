
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

__device__ __forceinline__ void ptk_body(const double i0, const double i1, double& o0) {
  const double t0 = (double)(((i0) + (i1)));
  o0 = (double)(t0);
}

struct EwDims { int ndim; long long shape[8]; long long st[3][8]; };

extern "C" __global__ void __launch_bounds__(256) ptk_ew_gen_3407c02cffde2853(const double* pi0, const double* pi1, double* po0, const EwDims dims, long long total) {
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
    long long rem = i;
    long long off0 = 0;
    long long off1 = 0;
    long long off2 = 0;
#pragma unroll
    for (int k = 8 - 1; k >= 0; --k) {
      if (k < dims.ndim) {
        const long long q = rem / dims.shape[k];
        const long long cidx = rem - q * dims.shape[k];
        rem = q;
        off0 += cidx * dims.st[0][k];
        off1 += cidx * dims.st[1][k];
        off2 += cidx * dims.st[2][k];
      }
    }
    double to0;
    ptk_body(pi0[off0], pi1[off1], to0);
    po0[off2] = to0;
  }
}
