
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

__device__ __forceinline__ void ptk_body(const double i0, const double i1, const double i2, double& o0, double& o1) {
  const double t0 = (double)(((0x1.0000000000000p-2) * (i1)));
  const double t1 = (double)(((0x1.0000000000000p-1) * (i2)));
  const double t2 = (double)(((t1) - (t0)));
  const double t3 = (double)(((t2) + (i0)));
  const double t4 = (double)(((-0x1.0000000000000p+0) * (t3) * (t3)));
  const double t5 = (double)(exp((double)(t4)));
  o0 = (double)(t3);
  o1 = (double)(t5);
}
struct ScDims { int ndim; long long shape[8]; long long st[3][8]; long long tstride[3]; long long store[2]; };
extern "C" __global__ void __launch_bounds__(256) ptk_scan_fused_3c294e691226d9b1(const double* __restrict__ pseq0, double* pst0, double* __restrict__ pnit0, const ScDims d, long long total, long long T) {
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gstride) {
    long long rem = e;
    long long off0 = 0;
    long long off1 = 0;
    long long off2 = 0;
#pragma unroll
    for (int k = 8 - 1; k >= 0; --k) {
      if (k < d.ndim) {
        const long long q = rem / d.shape[k]; const long long c = rem - q * d.shape[k]; rem = q;
        off0 += c * d.st[0][k];
        off1 += c * d.st[1][k];
        off2 += c * d.st[2][k];
      }
    }
    double w0_0 = pst0[off1 + 0LL * d.tstride[1]];
    double w0_1 = pst0[off1 + 1LL * d.tstride[1]];
    double sq0 = (T > 0) ? pseq0[off0] : (double)0;
    const int Ti = (int)T;
    int i0 = Ti;
    const int fs0 = (int)max(0LL, T - d.store[0]); i0 = min(i0, fs0);
    const int fn0 = (int)max(0LL, T - d.store[1]); i0 = min(i0, fn0);
    for (int i = 0; i < i0; ++i) {
      const double cur_sq0 = sq0;
      if (i + 1 < Ti) sq0 = pseq0[off0 + (long long)(i + 1) * d.tstride[0]];
      double nv0;
      double nn0;
      ptk_body(cur_sq0, w0_0, w0_1, nv0, nn0);
      w0_0 = w0_1;
      w0_1 = nv0;
    }
    const int st0 = (int)d.store[0];
    int sl0 = (int)((2LL + i0) % d.store[0]);
    long long wo0 = sl0 * d.tstride[1];
    const int stn0 = (int)d.store[1];
    int sn0 = stn0 > 0 ? (int)(i0 % d.store[1]) : 0;
    long long no0 = sn0 * d.tstride[2];
    for (int i = i0; i < Ti; ++i) {
      const double cur_sq0 = sq0;
      if (i + 1 < Ti) sq0 = pseq0[off0 + (long long)(i + 1) * d.tstride[0]];
      double nv0;
      double nn0;
      ptk_body(cur_sq0, w0_0, w0_1, nv0, nn0);
      w0_0 = w0_1;
      w0_1 = nv0;
      if (i >= fs0) pst0[off1 + wo0] = nv0;
      wo0 += d.tstride[1];
      if (++sl0 == st0) { sl0 = 0; wo0 = 0; }
      if (i >= fn0) pnit0[off2 + no0] = nn0;
      no0 += d.tstride[2];
      if (++sn0 == stn0) { sn0 = 0; no0 = 0; }
    }
  }
}