
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

__device__ __forceinline__ void ptk_body(const signed char i0, const long long i1, const long long i2, unsigned char& o0, unsigned char& o1, signed char& o2, long long& o3, float& o4, long long& o5, long long& o6, long long& o7, long long& o8, long long& o9, long long& o10, unsigned char& o11, long long& o12, float& o13, long long& o14, long long& o15) {
  const long long t0 = (long long)(ptk_floordiv<long long>((long long)(i1), (long long)(i2)));
  const long long t1 = (long long)(ptk_imod_py<long long>((long long)(i1), (long long)(i2)));
  const float t2 = (float)((((double)(i1)) / (i2)));
  const long long t3 = (long long)((((i2) > (i1)) ? (i2) : (i1)));
  const unsigned char t4 = (unsigned char)(((unsigned char)((i1) == (i2))));
  const long long t5 = (long long)(((i1) * (i1)));
  const long long t6 = (long long)((((i1) > 0) - ((i1) < 0)));
  const long long t7 = (long long)(((i1) & (i2)));
  const long long t8 = (long long)(((i1) | (i2)));
  const long long t9 = (long long)(((i1) ^ (i2)));
  const unsigned char t10 = (unsigned char)(((unsigned char)((i1) > (((signed char)0)))));
  const long long t11 = (long long)(((t10) ? (i1) : (i2)));
  const float t12 = (float)(((float)(i1)));
  const float t13 = (float)(((0x1.0000000000000p-1f) * (t12)));
  const long long t14 = (long long)(((i1) + (i0)));
  const signed char t15 = (signed char)((((i0) < 0) ? -(i0) : (i0)));
  const unsigned char t16 = (unsigned char)(((unsigned char)((i1) > (i2))));
  const unsigned char t17 = (unsigned char)(((unsigned char)(!(t16))));
  const unsigned char t18 = (unsigned char)(((unsigned char)((i0) > (((signed char)0)))));
  const unsigned char t19 = (unsigned char)(((t16) & (t18)));
  o0 = (unsigned char)(t19);
  o1 = (unsigned char)(t17);
  o2 = (signed char)(t15);
  o3 = (long long)(t14);
  o4 = (float)(t13);
  o5 = (long long)(t11);
  o6 = (long long)(t9);
  o7 = (long long)(t8);
  o8 = (long long)(t7);
  o9 = (long long)(t6);
  o10 = (long long)(t5);
  o11 = (unsigned char)(t4);
  o12 = (long long)(t3);
  o13 = (float)(t2);
  o14 = (long long)(t1);
  o15 = (long long)(t0);
}

#define VW 4
#define U 4
extern "C" __global__ void __launch_bounds__(256) ptk_ew_vec_2a8427eb3e72fc07(const signed char* __restrict__ pi0, const long long* __restrict__ pi1, const long long* __restrict__ pi2, unsigned char* __restrict__ po0, unsigned char* __restrict__ po1, signed char* __restrict__ po2, long long* __restrict__ po3, float* __restrict__ po4, long long* __restrict__ po5, long long* __restrict__ po6, long long* __restrict__ po7, long long* __restrict__ po8, long long* __restrict__ po9, long long* __restrict__ po10, unsigned char* __restrict__ po11, long long* __restrict__ po12, float* __restrict__ po13, long long* __restrict__ po14, long long* __restrict__ po15, long long rsi0, long long rsi1, long long rsi2, long long rso0, long long rso1, long long rso2, long long rso3, long long rso4, long long rso5, long long rso6, long long rso7, long long rso8, long long rso9, long long rso10, long long rso11, long long rso12, long long rso13, long long rso14, long long rso15, long long nchunks, unsigned int cpr, long long tail_start, long long n_total) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long base = gtid; base < nchunks; base += gstride * U) {
      PVec<signed char, VW> vi0[U];
      PVec<long long, VW> vi1[U];
      PVec<long long, VW> vi2[U];

#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          vi0[u] = ptk_ldv<signed char, VW>(pi0 + q * VW);
          vi1[u] = ptk_ldv<long long, VW>(pi1 + q * VW);
          vi2[u] = ptk_ldv<long long, VW>(pi2 + q * VW);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long q = base + (long long)u * gstride;
        if (q < nchunks) {

          PVec<unsigned char, VW> vo0;
          PVec<unsigned char, VW> vo1;
          PVec<signed char, VW> vo2;
          PVec<long long, VW> vo3;
          PVec<float, VW> vo4;
          PVec<long long, VW> vo5;
          PVec<long long, VW> vo6;
          PVec<long long, VW> vo7;
          PVec<long long, VW> vo8;
          PVec<long long, VW> vo9;
          PVec<long long, VW> vo10;
          PVec<unsigned char, VW> vo11;
          PVec<long long, VW> vo12;
          PVec<float, VW> vo13;
          PVec<long long, VW> vo14;
          PVec<long long, VW> vo15;
#pragma unroll
          for (int e = 0; e < VW; ++e) {
            ptk_body(vi0[u].v[e], vi1[u].v[e], vi2[u].v[e], vo0.v[e], vo1.v[e], vo2.v[e], vo3.v[e], vo4.v[e], vo5.v[e], vo6.v[e], vo7.v[e], vo8.v[e], vo9.v[e], vo10.v[e], vo11.v[e], vo12.v[e], vo13.v[e], vo14.v[e], vo15.v[e]);
          }
          ptk_stv<unsigned char, VW>(po0 + q * VW, vo0);
          ptk_stv<unsigned char, VW>(po1 + q * VW, vo1);
          ptk_stv<signed char, VW>(po2 + q * VW, vo2);
          ptk_stv<long long, VW>(po3 + q * VW, vo3);
          ptk_stv<float, VW>(po4 + q * VW, vo4);
          ptk_stv<long long, VW>(po5 + q * VW, vo5);
          ptk_stv<long long, VW>(po6 + q * VW, vo6);
          ptk_stv<long long, VW>(po7 + q * VW, vo7);
          ptk_stv<long long, VW>(po8 + q * VW, vo8);
          ptk_stv<long long, VW>(po9 + q * VW, vo9);
          ptk_stv<long long, VW>(po10 + q * VW, vo10);
          ptk_stv<unsigned char, VW>(po11 + q * VW, vo11);
          ptk_stv<long long, VW>(po12 + q * VW, vo12);
          ptk_stv<float, VW>(po13 + q * VW, vo13);
          ptk_stv<long long, VW>(po14 + q * VW, vo14);
          ptk_stv<long long, VW>(po15 + q * VW, vo15);
        }
      }
  }
  // flat tail (rows == 1 only): the last n_total % VW elements
  for (long long i = tail_start + gtid; i < n_total; i += gstride) {
      unsigned char to0;
      unsigned char to1;
      signed char to2;
      long long to3;
      float to4;
      long long to5;
      long long to6;
      long long to7;
      long long to8;
      long long to9;
      long long to10;
      unsigned char to11;
      long long to12;
      float to13;
      long long to14;
      long long to15;
      ptk_body(pi0[i], pi1[i], pi2[i], to0, to1, to2, to3, to4, to5, to6, to7, to8, to9, to10, to11, to12, to13, to14, to15);
      po0[i] = to0;
      po1[i] = to1;
      po2[i] = to2;
      po3[i] = to3;
      po4[i] = to4;
      po5[i] = to5;
      po6[i] = to6;
      po7[i] = to7;
      po8[i] = to8;
      po9[i] = to9;
      po10[i] = to10;
      po11[i] = to11;
      po12[i] = to12;
      po13[i] = to13;
      po14[i] = to14;
      po15[i] = to15;
  }
}
