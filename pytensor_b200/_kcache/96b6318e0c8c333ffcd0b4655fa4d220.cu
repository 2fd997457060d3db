
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

__device__ __forceinline__ void ptk_body(const double i0, const double i1, const double i2, const double i3, const double i4, const double i5, const double i6, const double i7, const double i8, double& o0, double& o1, double& o2) {
  const double t0 = (double)(((-0x1.0000000000000p+0) * (i7) * (i8)));
  const double t1 = (double)(((-0x1.0000000000000p-1) * (i4)));
  const double t2 = (double)(((t1) + (t1) + (i6) + (t0)));
  const double t3 = (double)((-(i5)));
  const double t4 = (double)(((-0x1.0000000000000p-1) * (i3)));
  const double t5 = (double)(((t4) + (t4) + (t3)));
  const double t6 = (double)(((0x1.0000000000000p-1) * (i4) * (i4)));
  const double t7 = (double)(((0x1.0000000000000p-1) * (i3) * (i3)));
  const double t8 = (double)(((t7) + (t6)));
  const double t9 = (double)(((-0x1.d67f1c864beb4p+0) + (i0) + (i1) + (i2)));
  const double t10 = (double)(((t9) - (t8)));
  o0 = (double)(t10);
  o1 = (double)(t5);
  o2 = (double)(t2);
}

#define VW 4
#define U 4
extern "C" __global__ void __launch_bounds__(256) ptk_ew_vec_04ee14f5741d0ee5(const double* pi0, const double* pi1, const double* pi2, const double* pi3, const double* pi4, const double* pi5, const double* pi6, const double* pi7, const double* pi8, double* po0, double* po1, double* po2, long long rsi0, long long rsi1, long long rsi2, long long rsi3, long long rsi4, long long rsi5, long long rsi6, long long rsi7, long long rsi8, long long rso0, long long rso1, long long rso2, long long nchunks, unsigned int cpr, long long tail_start, long long n_total) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long base = gtid; base < nchunks; base += gstride * U) {
      PVec<double, VW> vi0[U];
      PVec<double, VW> vi1[U];
      PVec<double, VW> vi2[U];
      PVec<double, VW> vi3[U];
      PVec<double, VW> vi4[U];
      PVec<double, VW> vi5[U];
      double vi6[U];
      PVec<double, VW> vi7[U];
      PVec<double, VW> vi8[U];

#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          vi0[u] = ptk_ldv<double, VW>(pi0 + q * VW);
          vi1[u] = ptk_ldv<double, VW>(pi1 + q * VW);
          vi2[u] = ptk_ldv<double, VW>(pi2 + q * VW);
          vi3[u] = ptk_ldv<double, VW>(pi3 + q * VW);
          vi4[u] = ptk_ldv<double, VW>(pi4 + q * VW);
          vi5[u] = ptk_ldv<double, VW>(pi5 + q * VW);
          vi6[u] = pi6[0];
          vi7[u] = ptk_ldv<double, VW>(pi7 + q * VW);
          vi8[u] = ptk_ldv<double, VW>(pi8 + q * VW);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          PVec<double, VW> vo0;
          PVec<double, VW> vo1;
          PVec<double, VW> vo2;
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            ptk_body(vi0[u].v[e], vi1[u].v[e], vi2[u].v[e], vi3[u].v[e], vi4[u].v[e], vi5[u].v[e], vi6[u], vi7[u].v[e], vi8[u].v[e], vo0.v[e], vo1.v[e], vo2.v[e]);
          }
          ptk_stv<double, VW>(po0 + q * VW, vo0);
          ptk_stv<double, VW>(po1 + q * VW, vo1);
          ptk_stv<double, VW>(po2 + q * VW, vo2);
        }
      }
  }
  // flat tail (rows == 1 only): the last n_total % VW elements
  for (long long i = tail_start + gtid; i < n_total; i += gstride) {
      double to0;
      double to1;
      double to2;
      ptk_body(pi0[i], pi1[i], pi2[i], pi3[i], pi4[i], pi5[i], pi6[0], pi7[i], pi8[i], to0, to1, to2);
      po0[i] = to0;
      po1[i] = to1;
      po2[i] = to2;
  }
}
