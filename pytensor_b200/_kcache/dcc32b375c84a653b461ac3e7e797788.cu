
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


template <typename T, int N> struct __align__(sizeof(T) * N) PVec { T v[N]; };
template <typename T, int N> __device__ __forceinline__ PVec<T, N> ptk_ldv(const T* p) {
  return *reinterpret_cast<const PVec<T, N>*>(p);
}
template <typename T, int N> __device__ __forceinline__ void ptk_stv(T* p, const PVec<T, N>& v) {
  *reinterpret_cast<PVec<T, N>*>(p) = v;
}

__device__ __forceinline__ void ptk_body(const double i0, const double i1, double& o0) {
  const double t0 = (double)(((i0) + (i1)));
  o0 = (double)(t0);
}

#define VW 4
#define U 4
extern "C" __global__ void __launch_bounds__(256) ptk_ew_vec_5584f1b38315f318(const double* pi0, const double* pi1, double* po0, long long rsi0, long long rsi1, long long rso0, long long nchunks, unsigned int cpr, long long tail_start, long long n_total) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long base = gtid; base < nchunks; base += gstride * U) {
      double vi0[U];
      double vi1[U];

#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          vi0[u] = pi0[0];
          vi1[u] = pi1[0];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          PVec<double, VW> vo0;
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            ptk_body(vi0[u], vi1[u], vo0.v[e]);
          }
          ptk_stv<double, VW>(po0 + q * VW, vo0);
        }
      }
  }
  // flat tail (rows == 1 only): the last n_total % VW elements
  for (long long i = tail_start + gtid; i < n_total; i += gstride) {
      double to0;
      ptk_body(pi0[0], pi1[0], to0);
      po0[i] = to0;
  }
}
