// Random draws on the device (reference: RandomVariable, pytensor/tensor/random/op.py:49, whose perform
// :457-468 calls numpy.random.Generator methods on the host).  numpy's PCG64 stream with its rejection samplers is inherently
// sequential; here every output element owns a COUNTER-BASED stream instead: Philox4x32-10 keyed by 128 bits the caller
// takes from the host Generator (which thereby advances), counter = (element index, draw round).  Same seed => same
// draws, any element computable independently, no state in device memory.  Values therefore differ from the C linker's;
// parity is distributional (moments + Kolmogorov-Smirnov in tests/test_gpu_random.py), shapes / dtypes / the
// advance-the-generator contract are exact.
#include <math.h>
#include <algorithm>
#include "ptk_common.h"

namespace {

using ptk::fail;

struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t ka, uint32_t kb) const {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ ka, n1 = lo1, n2 = hi0 ^ c[3] ^ kb, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  __device__ __forceinline__ void block(uint32_t (&c)[4]) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round(c, ka, kb);
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
  }
};

// 53-bit uniforms in (0, 1): never 0, never 1 (safe for log / tan)
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
  const uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double ptk_tanpi(double t) { return sinpi(t) / cospi(t); }

struct Draws {   // counter-based supply of uniforms / normals for ONE output element
  Philox ph;
  uint64_t idx, seed_hi;
  uint32_t round_;
  uint32_t c[4];
  int have;
  __device__ __forceinline__ Draws(uint64_t k, uint64_t s, uint64_t i) : idx(i), seed_hi(s), round_(0), have(0) {
    ph.k0 = (uint32_t)k;
    ph.k1 = (uint32_t)(k >> 32);
  }
  __device__ __forceinline__ double uniform() {
    if (have == 0) {
      c[0] = (uint32_t)idx; c[1] = (uint32_t)(idx >> 32); c[2] = round_++ ^ (uint32_t)seed_hi; c[3] = (uint32_t)(seed_hi >> 32);
      ph.block(c);
      have = 2;
    }
    --have;
    return have == 1 ? u01(c[0], c[1]) : u01(c[2], c[3]);
  }
  __device__ __forceinline__ double normal() {   // Box-Muller, one of the pair
    const double u = uniform(), v = uniform();
    return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
  }
  __device__ double gamma(double a) {            // Marsaglia & Tsang (2000); a < 1 through the a+1 boost
    if (!(a > 0.0)) return a == 0.0 ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
    double boost = 1.0;
    if (a < 1.0) {
      boost = pow(uniform(), 1.0 / a);
      a += 1.0;
    }
    const double d = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * d);
    for (int it = 0; it < 64; ++it) {
      const double x = normal();
      double v = 1.0 + cc * x;
      if (v <= 0.0) continue;
      v = v * v * v;
      const double u = uniform();
      if (u < 1.0 - 0.0331 * (x * x) * (x * x) || log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
    }
    return boost * d;
  }
};

enum Dist { UNIFORM = 0, NORMAL = 1, HALFNORMAL = 2, LOGNORMAL = 3, EXPONENTIAL = 4, LAPLACE = 5, LOGISTIC = 6, GUMBEL = 7,
            CAUCHY = 8, BERNOULLI = 9, GAMMA = 10, BETA = 11, INTEGERS = 12, WEIBULL = 13, PARETO = 14, HALFCAUCHY = 15,
            INVGAMMA = 16, STUDENTT = 17, N_DIST = 18 };

template <typename OUT>
__global__ void __launch_bounds__(256) random_kernel(int dist, OUT* __restrict__ out, long long n, uint64_t key, uint64_t seed,
                                                     const double* __restrict__ p0, long long s0,
                                                     const double* __restrict__ p1, long long s1,
                                                     const double* __restrict__ p2, long long s2) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    Draws g(key, seed, (uint64_t)i);
    const double a = p0 ? p0[i * s0] : 0.0, b = p1 ? p1[i * s1] : 1.0, c = p2 ? p2[i * s2] : 1.0;
    double x;
    switch (dist) {
      case UNIFORM: x = a + (b - a) * g.uniform(); break;
      case NORMAL: x = a + b * g.normal(); break;
      case HALFNORMAL: x = a + b * fabs(g.normal()); break;
      case LOGNORMAL: x = exp(a + b * g.normal()); break;
      case EXPONENTIAL: x = -a * log(g.uniform()); break;                       // a = scale
      case LAPLACE: { const double u = g.uniform() - 0.5; x = a - b * copysign(log(1.0 - 2.0 * fabs(u)), u); } break;
      case LOGISTIC: { const double u = g.uniform(); x = a + b * log(u / (1.0 - u)); } break;
      case GUMBEL: x = a - b * log(-log(g.uniform())); break;
      case CAUCHY: x = a + b * ptk_tanpi(g.uniform() - 0.5); break;
      case HALFCAUCHY: x = a + b * fabs(ptk_tanpi(g.uniform() - 0.5)); break;
      case BERNOULLI: x = g.uniform() < a ? 1.0 : 0.0; break;                    // a = p
      case GAMMA: x = b * g.gamma(a); break;                                     // a = shape, b = scale
      case INVGAMMA: x = b / g.gamma(a); break;                                  // a = shape, b = scale
      case BETA: { const double ga = g.gamma(a), gb = g.gamma(b); x = ga / (ga + gb); } break;
      case INTEGERS: x = floor(a + (b - a) * g.uniform()); if (x >= b) x = b - 1.0; break;   // [low, high)
      case WEIBULL: x = pow(-log(g.uniform()), 1.0 / a); break;                  // a = shape
      case PARETO: x = b * pow(g.uniform(), -1.0 / a); break;                    // scipy's form (x >= scale), a = shape, b = scale
      case STUDENTT: { const double z = g.normal(), ch = 2.0 * g.gamma(0.5 * a); x = b + c * z / sqrt(ch / a); } break;  // a=df
      default: x = 0.0;
    }
    out[i] = (OUT)x;
  }
}

}  // namespace

extern "C" ptk_status ptk_random_fill(int dist, int dtype, void* out, int64_t n, uint64_t key, uint64_t seed, const void* p0,
                                      int64_t s0, const void* p1, int64_t s1, const void* p2, int64_t s2, void* stream) {
  PTK_REQUIRE_INIT();
  if (dist < 0 || dist >= N_DIST) return fail(PTK_ERR_ARG, "ptk_random_fill: unknown distribution");
  if (n <= 0) return PTK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned g = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ptk::sm_count() * 16));
#define PTK_RND(T) random_kernel<T><<<g, 256, 0, st>>>(dist, (T*)out, n, key, seed, (const double*)p0, s0, (const double*)p1, s1, (const double*)p2, s2); break;
  switch (dtype) {
    case PTK_F32: PTK_RND(float)
    case PTK_F64: PTK_RND(double)
    case PTK_I64: PTK_RND(int64_t)
    case PTK_I32: PTK_RND(int32_t)
    case PTK_I16: PTK_RND(int16_t)
    case PTK_I8: PTK_RND(int8_t)
    case PTK_U8: case PTK_BOOL: PTK_RND(uint8_t)
    default: return fail(PTK_ERR_UNSUPPORTED, "ptk_random_fill: output dtype");
  }
#undef PTK_RND
  PTK_LAUNCH_CHECK("random_fill");
  return PTK_OK;
}
