// Dense factor / triangular-solve kernels for Cholesky and SolveTriangular
// (pytensor/tensor/linalg/decomposition/cholesky.py:18 potrf :52-83; solvers/triangular.py:13 trtrs :41-71).
// Warp-cooperative, one CTA per matrix (potrf) or per (matrix, 32-wide RHS panel) (trsm); row-major storage.
// Failure is signalled the reference's way: NaN-fill, never an exception (cholesky.py:78-80, triangular.py:68-69).
#include <math_constants.h>
#include "ptk_common.h"

namespace {

template <typename T> __device__ __forceinline__ T nan_of();
template <> __device__ __forceinline__ float nan_of<float>() { return CUDART_NAN_F; }
template <> __device__ __forceinline__ double nan_of<double>() { return CUDART_NAN; }

// Left-looking Cholesky. LOWER: element (i,j) lives at A[i*n+j]; otherwise the factor is U = L^T and (i,j) of L lives
// at A[j*n+i]. Column j: every warp owns rows i>j and forms dot(L[i,:j], L[j,:j]) with lanes along k.
template <typename T, bool LOWER>
__global__ void __launch_bounds__(512) potrf_kernel(T* __restrict__ Aall, int64_t n) {
  T* A = Aall + (int64_t)blockIdx.x * n * n;
  extern __shared__ unsigned char smem_raw[];
  T* rowj = reinterpret_cast<T*>(smem_raw);  // L[j, 0..j)
  __shared__ T s_d;
  __shared__ int s_bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  auto at = [&](int64_t i, int64_t j) -> T& { return LOWER ? A[i * n + j] : A[j * n + i]; };
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (int64_t j = 0; j < n; ++j) {
    for (int64_t k = tid; k < j; k += blockDim.x) rowj[k] = at(j, k);
    __syncthreads();
    if (warp == 0) {
      T s = T(0);
      for (int64_t k = lane; k < j; k += 32) s += rowj[k] * rowj[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) {
        T d = at(j, j) - s;
        if (!(d > T(0))) s_bad = 1;
        d = sqrt(d);
        at(j, j) = d;
        s_d = d;
      }
    }
    __syncthreads();
    if (s_bad) break;
    const T inv = T(1) / s_d;
    for (int64_t i = j + 1 + warp; i < n; i += nwarps) {
      T s = T(0);
      for (int64_t k = lane; k < j; k += 32) s += at(i, k) * rowj[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) at(i, j) = (at(i, j) - s) * inv;
    }
    __syncthreads();
  }
  const bool bad = s_bad != 0;
  const T nanv = nan_of<T>();
  for (int64_t e = tid; e < n * n; e += blockDim.x) {
    int64_t r = e / n, c = e - r * n;
    if (bad) A[e] = nanv;
    else if (LOWER ? (c > r) : (c < r)) A[e] = T(0);  // clean=True: zero the other triangle
  }
}

// Substitution for op(A) X = B, one CTA per (matrix, panel of 32 RHS columns); 32 columns x 8 k-groups.
template <typename T>
__global__ void __launch_bounds__(256) trsm_kernel(const T* __restrict__ Aall, T* __restrict__ Ball, int64_t n,
                                                   int64_t nrhs, int lower, int trans, int unit_diag) {
  const T* A = Aall + (int64_t)blockIdx.y * n * n;
  T* B = Ball + (int64_t)blockIdx.y * n * nrhs;
  __shared__ T red[8][33];
  __shared__ int s_sing;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + lane;
  const bool active = col < nrhs;
  const bool fwd = (lower != 0) != (trans != 0);
  auto opA = [&](int64_t i, int64_t k) -> T { return trans ? A[k * n + i] : A[i * n + k]; };
  if (threadIdx.x == 0) s_sing = 0;
  __syncthreads();
  if (!unit_diag) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x)
      if (A[i * n + i] == T(0)) s_sing = 1;
  }
  __syncthreads();
  if (s_sing) {
    const T nanv = nan_of<T>();
    for (int64_t i = grp; i < n; i += 8)
      if (active) B[i * nrhs + col] = nanv;
    return;
  }
  for (int64_t step = 0; step < n; ++step) {
    const int64_t i = fwd ? step : (n - 1 - step);
    T s = T(0);
    if (active) {
      if (fwd) for (int64_t k = grp; k < i; k += 8) s += opA(i, k) * B[k * nrhs + col];
      else     for (int64_t k = i + 1 + grp; k < n; k += 8) s += opA(i, k) * B[k * nrhs + col];
    }
    red[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && active) {
      T tot = T(0);
#pragma unroll
      for (int g = 0; g < 8; ++g) tot += red[g][lane];
      T v = B[i * nrhs + col] - tot;
      if (!unit_diag) v = v / A[i * n + i];
      B[i * nrhs + col] = v;
    }
    __syncthreads();
  }
}

}  // namespace

using namespace ptk;

extern "C" {

ptk_status ptk_potrf(int dtype, void* A, int64_t n, int64_t batch, int lower, void* stream) {
  PTK_REQUIRE_INIT();
  if (n == 0 || batch == 0) return PTK_OK;
  if (batch > 2147483647LL) return fail(PTK_ERR_ARG, "ptk_potrf: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)n * dtype_size(dtype);
  if (smem > 200 * 1024) return fail(PTK_ERR_UNSUPPORTED, "ptk_potrf: n too large for the single-CTA panel kernel");
  int threads = n >= 256 ? 512 : (n >= 64 ? 256 : 64);
#define PTK_POTRF(T, L)                                                                                        \
  do {                                                                                                         \
    PTK_CUDA(cudaFuncSetAttribute(potrf_kernel<T, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    potrf_kernel<T, L><<<(unsigned)batch, threads, smem, st>>>((T*)A, n);                                     \
  } while (0)
  if (dtype == PTK_F32) { if (lower) PTK_POTRF(float, true); else PTK_POTRF(float, false); }
  else if (dtype == PTK_F64) { if (lower) PTK_POTRF(double, true); else PTK_POTRF(double, false); }
  else return fail(PTK_ERR_UNSUPPORTED, "ptk_potrf: dtype must be float32 or float64");
#undef PTK_POTRF
  PTK_LAUNCH_CHECK("potrf");
  return PTK_OK;
}

ptk_status ptk_trsm(int dtype, const void* A, void* B, int64_t n, int64_t nrhs, int64_t batch, int lower, int trans,
                    int unit_diag, void* stream) {
  PTK_REQUIRE_INIT();
  if (n == 0 || nrhs == 0 || batch == 0) return PTK_OK;
  if (batch > 65535) return fail(PTK_ERR_ARG, "ptk_trsm: batch > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((nrhs + 31) / 32), (unsigned)batch);
  if (dtype == PTK_F32)
    trsm_kernel<float><<<grid, 256, 0, st>>>((const float*)A, (float*)B, n, nrhs, lower, trans, unit_diag);
  else if (dtype == PTK_F64)
    trsm_kernel<double><<<grid, 256, 0, st>>>((const double*)A, (double*)B, n, nrhs, lower, trans, unit_diag);
  else return fail(PTK_ERR_UNSUPPORTED, "ptk_trsm: dtype must be float32 or float64");
  PTK_LAUNCH_CHECK("trsm");
  return PTK_OK;
}

}  // extern "C"
