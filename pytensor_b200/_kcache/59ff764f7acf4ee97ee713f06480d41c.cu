
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


template <typename T> __device__ __forceinline__ T ptk_nanmax(T a, T b) { return (b > a) ? b : ((a >= b) ? a : (a + b)); }
template <typename T> __device__ __forceinline__ T ptk_nanmin(T a, T b) { return (b < a) ? b : ((a <= b) ? a : (a + b)); }
template <typename T> __device__ __forceinline__ T ptk_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <> __device__ __forceinline__ unsigned char ptk_shfl_xor<unsigned char>(unsigned char v, int m) { return (unsigned char)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ signed char ptk_shfl_xor<signed char>(signed char v, int m) { return (signed char)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ short ptk_shfl_xor<short>(short v, int m) { return (short)__shfl_xor_sync(0xffffffffu, (int)v, m); }
template <> __device__ __forceinline__ unsigned short ptk_shfl_xor<unsigned short>(unsigned short v, int m) { return (unsigned short)__shfl_xor_sync(0xffffffffu, (int)v, m); }

__device__ __forceinline__ void ptk_body(const double i0, double& o0) {
  const double t0 = (double)((i0));
  o0 = (double)(t0);
}
typedef double ACC;
typedef double OUT;
__device__ __forceinline__ ACC ptk_red(ACC a, ACC b) { return (ACC)ptk_nanmax((a), (b)); }
#define VW 1
#define TPR 32

extern "C" __global__ void __launch_bounds__(256) ptk_red_row_0f186fc7aa55cc7c(const double* __restrict__ pi0, void* __restrict__ pred, long long rsi0, long long rows, long long cols, int nsplit) {
  const int lane_in_row = threadIdx.x % TPR;
  const int row_in_block = threadIdx.x / TPR;
  const long long ncv = cols / VW;                       // vector chunks per row
  const long long per_split = (ncv + nsplit - 1) / nsplit;
  const int split = blockIdx.y;
  const long long cv_lo = (long long)split * per_split;
  const long long cv_hi = (cv_lo + per_split < ncv) ? (cv_lo + per_split) : ncv;
  for (long long rb = (long long)blockIdx.x * 8; rb < rows; rb += (long long)gridDim.x * 8) {
    const long long r = rb + row_in_block;
    ACC acc = (ACC)__longlong_as_double(0xfff0000000000000LL);
    if (r < rows) {
      for (long long cv = cv_lo + lane_in_row; cv < cv_hi; cv += TPR) {
        const long long c = cv * VW;
        const PVec<double, VW> vi0 = ptk_ldv<double, VW>(pi0 + r * rsi0 + c);
        PVec<double, VW> vo0;
        #pragma unroll
        for (int e = 0; e < VW; ++e) {
          ptk_body(vi0.v[e], vo0.v[e]);
          acc = ptk_red(acc, (ACC)vo0.v[e]);
        }

      }
      if (split == nsplit - 1) {
        for (long long c = ncv * VW + lane_in_row; c < cols; c += TPR) {
        double to0;
          ptk_body(pi0[r * rsi0 + c], to0);
          acc = ptk_red(acc, (ACC)to0);

        }
      }
    }

    #pragma unroll
    for (int m = 16; m > 0; m >>= 1) acc = ptk_red(acc, ptk_shfl_xor<ACC>(acc, m));

    if (lane_in_row == 0 && r < rows) {
      if (nsplit == 1) reinterpret_cast<OUT*>(pred)[r] = (OUT)acc;
      else reinterpret_cast<ACC*>(pred)[r * nsplit + split] = acc;
    }
  }
}
