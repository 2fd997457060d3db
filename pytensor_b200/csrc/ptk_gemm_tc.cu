// bf16 tensor-core GEMM for fp32 graphs on sm_100a: tcgen05.mma (cta_group::1, M=128, N=256, K=16) with fp32
// accumulators in TMEM, operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a 4-stage mbarrier ring,
// persistent CTAs (one per SM) with a double-buffered TMEM accumulator so that the epilogue of tile i overlaps the MMAs
// of tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM allocator,
// warps 4-7 = epilogue (tcgen05.ld 32x32b -> alpha/beta/bias/tanh -> global).
//
// Replaces the sgemm_ call the C linker emits for Dot22 / Gemm (pytensor/tensor/blas/c_code/codegen.py:463-540) on the
// "bf16 on tcgen05" configuration of BASELINE.json: the graph dtype stays float32 (the reference has no bfloat16 dtype,
// pytensor/tensor/type.py:40-55), operands are rounded to bf16 by a conversion pass into a caller-owned workspace
// (A as [M,K] K-major, B transposed to [N,K] K-major), products accumulate in fp32.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <algorithm>
#include "ptk_common.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;  // 2 accumulators x 256 fp32 columns
constexpr int NUM_THREADS = 256;
constexpr uint32_t A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr uint32_t B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;  // 32 KiB
constexpr uint32_t STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// ---- PTX wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* smem_dst, int32_t c0, int32_t c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//  [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//  [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between row groups) | [46,48) version = 1 |
//  [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10, both K-major,
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct EpiParams {
  float alpha, beta;
  float* C;
  long long sc0, sc1;
  const float* bias;
  int act;
  int M, N, K;
  __nv_bfloat16* Cbf;  // optional bf16 copy of the result, row-major [M, ldcbf] (the next layer's K-major A operand)
  long long ldcbf;
  int out_pieces;      // pair kernel: 1 = Cbf is a plain bf16 copy; 3 = the three-piece split of the result (piece i at rows
  long long cbf_rows;  // [i * cbf_rows, ...) of Cbf) — the staged A operand of a following fp32-accurate product
  int split_tail;      // pair kernel: split the units of the last partial round into 256 x 128 halves (PTK_GEMM_SPLIT=0: off)
  // fp32-accurate mode (pair kernel): every fp32 operand is staged as THREE bf16 pieces x = x1 + x2 + x3 (8 mantissa bits
  // each) stacked along the rows of the staging matrix (piece i of A at rows [i * a_rows, ...)); per k-block the kernel
  // accumulates `terms` piece products into the same fp32 TMEM accumulator, smallest first:
  //   terms = 6: A3B1 + A2B2 + A1B3 + A2B1 + A1B2 + A1B1   (drops only O(2^-24) terms: below sgemm's own rounding noise)
  //   terms = 3:                      A2B1 + A1B2 + A1B1   (~4e-6 of the output scale at K = 4096)
  //   terms = 1: plain bf16 operands (the CUDA_BF16 mode)
  int terms;
  int a_rows, b_rows;  // row pitch between the stacked pieces
  // pair kernel: k-blocks per ACCUMULATION CHUNK (0 = the whole K in one chunk).  The tensor core adds into its fp32
  // accumulator with truncation, a bias that grows linearly with the length of the accumulation chain (measured: ~6e-9 x K
  // of the output scale, 2.4e-5 at K = 4096 — more than the whole 1e-5 budget of the fp32-accurate mode).  With chunks the
  // TMEM accumulator only ever holds the sum over `kchunk` k-blocks; the epilogue adds every chunk into C in global memory
  // with round-to-nearest fp32 adds (the same thread owns the same elements for all chunks of a tile, so the
  // read-modify-write needs no synchronisation) and applies bias / activation / the bf16 copy after the last chunk.
  int kchunk;
  // pair kernel, fp32-accurate mode with an error-free leading piece (see row_absmax_kernel): the A1 x B1 products go
  // to the accumulator at TMEM columns [0, 256), the correction products to the one at [256, 512); the epilogue adds the
  // two (round to nearest).  Both buffers form ONE accumulator stage, so the epilogue of a chunk does not overlap the
  // next chunk's MMAs.
  int exact_main;
  // staged output pieces (Cbf, out_pieces == 3): leading piece aligned to the fixed exponent out_exp (operands known to lie
  // in [-1, 1]) so that it can feed an exact-main product; PTK_NO_EXP: ordinary bf16 split
  int out_exp;
  float out_scale, out_inv;   // 2^out_exp and 2^-out_exp (exact powers of two: the alignment is two multiplies and a rint)
};
#define PTK_NO_EXP (-100000)
// piece indices of the term sequence; a run of `terms` entries ending at index 5 is used
__device__ __constant__ int kPieceA[6] = {2, 1, 0, 1, 0, 0};
__device__ __constant__ int kPieceB[6] = {0, 1, 2, 0, 1, 0};

// kCluster == 2: CTA pairs (cluster 2x1x1) share the B tile of a 256-row super-tile — each CTA loads HALF of it and
// TMA-multicasts that half into both CTAs' shared memory, cutting the L2->SM traffic per CTA from 48 to 32 KiB per
// k-block (the single-CTA kernel is L2-bandwidth bound: 1.6 GB per 4096^3 GEMM at ~12 TB/s).  A shared-memory stage
// may be refilled only when BOTH CTAs' MMAs have drained it: the stage's "empty" barrier counts 2 arrivals and every
// tcgen05.commit is multicast to the pair.
template <int kCluster>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                              // STAGES x 16 KiB
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;     // STAGES x 32 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [ACC_STAGES]
  uint64_t* tmem_empty = bars + 2 * STAGES + ACC_STAGES;  // [ACC_STAGES]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M, n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const uint32_t crank = (kCluster > 1) ? cluster_ctarank() : 0u;
  // work decomposition: a "unit" is kCluster vertically adjacent M-tiles of one N-tile; units are dealt round-robin
  const int m_groups = (m_tiles + kCluster - 1) / kCluster;
  const int num_units = m_groups * n_tiles;
  const int unit0 = blockIdx.x / kCluster, unit_stride = gridDim.x / kCluster;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kCluster);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // the peer's barriers must be initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = unit0; u < num_units; u += unit_stride) {
        const int tm = (u % m_groups) * kCluster + (int)crank, tn = u / m_groups;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, tm * BLOCK_M);
          if (kCluster == 1) {
            tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, kb * BLOCK_K, tn * BLOCK_N);
          } else {
            constexpr uint32_t HALF = B_STAGE_BYTES / 2;  // 128 rows of the B tile
            tma_load_2d_mc(&tmap_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + crank * HALF, kb * BLOCK_K,
                           tn * BLOCK_N + (int)crank * (BLOCK_N / 2), (uint16_t)0x3);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = unit0; u < num_units; u += unit_stride) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t a_desc = make_smem_desc(a_addr + k * UMMA_K * 2);
            const uint64_t b_desc = make_smem_desc(b_addr + k * UMMA_K * 2);
            umma_f16(d_tmem, a_desc, b_desc, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot once these MMAs have read it (in both CTAs of the pair when the B tile is shared)
          if (kCluster == 1) umma_commit(&empty_bar[stage]);
          else umma_commit_mc(&empty_bar[stage], (uint16_t)0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = unit0; u < num_units; u += unit_stride) {
      const int tm = (u % m_groups) * kCluster + (int)crank, tn = u / m_groups;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long row = (long long)tm * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      float* crow = p.C + row * p.sc0;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
        const long long col0 = (long long)tn * BLOCK_N + c0;
        if (row_ok && col0 < p.N) {
          const bool full = (col0 + 32 <= p.N);
          if (full && p.sc1 == 1 && p.beta == 0.0f && ((((uintptr_t)(crow + col0)) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 v;
              float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x = p.alpha * __uint_as_float(r[j + e]);
                if (p.bias) x += p.bias[col0 + j + e];
                if (p.act == 1) x = tanhf(x);
                vv[e] = x;
              }
              *reinterpret_cast<float4*>(crow + col0 + j) = v;
              if (p.Cbf) {
                __nv_bfloat162 lo = __floats2bfloat162_rn(vv[0], vv[1]), hi = __floats2bfloat162_rn(vv[2], vv[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(p.Cbf + row * p.ldcbf + col0 + j) = pk;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const long long col = col0 + j;
              if (col < p.N) {
                float* dst = crow + col * p.sc1;
                float x = p.alpha * __uint_as_float(r[j]);
                if (p.beta != 0.0f) x += p.beta * (*dst);
                if (p.bias) x += p.bias[col];
                if (p.act == 1) x = tanhf(x);
                *dst = x;
                if (p.Cbf) p.Cbf[row * p.ldcbf + col] = __float2bfloat16_rn(x);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // nobody leaves while the peer may still multicast into / signal this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ====================================================================================================================
// cta_group::2 ("pair") kernel: two CTAs of a 2x1x1 cluster (one TPC) run ONE 256 x 256 UMMA tile.  Each CTA stages only
// its own 128 rows of A and HALF of the B tile (128 of the 256 N-rows), so per k-block a CTA writes 32 KiB into shared
// memory instead of 48 KiB and the MMA reads 8 KiB instead of 12 KiB per k16 step: shared-memory traffic drops from
// ~192 B/clk (above the 128 B/clk port, the limiter of the single-CTA kernel at ~60-67 % tensor-pipe) to ~128 B/clk, and
// the ring deepens from 4 to 6 stages.  Only the leader CTA (cluster rank 0) issues tcgen05.mma; both CTAs' TMA loads
// complete on the LEADER's "full" barrier; tcgen05.commit is multicast to both CTAs ("empty" + accumulator-ready
// barriers); both CTAs' epilogue warps arrive on the leader's "accumulator drained" barrier.
// ====================================================================================================================
constexpr int P_STAGES = 6;
constexpr uint32_t P_A_BYTES = BLOCK_M * BLOCK_K * 2;         // 16 KiB: this CTA's 128 rows of A
constexpr uint32_t P_B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;   // 16 KiB: this CTA's half of the B tile
constexpr uint32_t P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;     // 32 KiB
constexpr int EPI_PITCH = 36;                                  // floats per staged row: 16-byte aligned, conflict-free
constexpr uint32_t P_EPI_BYTES = 4 * 32 * EPI_PITCH * 4;       // one 32 x 32 fp32 transposition tile per epilogue warp
constexpr uint32_t P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + 1024 + 256 + P_EPI_BYTES;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                // shared::cluster address of the same offset in CTA rank 0

__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* smem_dst, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// kExact: compile-time copy of EpiParams::exact_main — the plain instantiation keeps the double-buffered accumulator
// bookkeeping and the single-accumulator epilogue as constants (the bf16 mode's inner loops carry none of the
// exact mode's selects).
template <bool kExact>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tc_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_bh, EpiParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                // P_STAGES x 16 KiB
  uint8_t* smem_b = smem + P_STAGES * P_A_BYTES;         // P_STAGES x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full_bar = bars;                             // [P_STAGES]   (used in the leader only)
  uint64_t* empty_bar = bars + P_STAGES;                 // [P_STAGES]   (one per CTA, signalled by the leader's commits)
  uint64_t* tmem_full = bars + 2 * P_STAGES;             // [ACC_STAGES] (one per CTA)
  uint64_t* tmem_empty = bars + 2 * P_STAGES + ACC_STAGES;  // [ACC_STAGES] (leader only; 8 arrivals = 4 warps x 2 CTAs)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * P_STAGES + 2 * ACC_STAGES);
  float* epi_tiles = reinterpret_cast<float*>(smem + P_STAGES * P_STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);  // 256-row tiles
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_units = m_tiles * n_tiles;
  const int k_blocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kchunk = (p.kchunk > 0 && p.kchunk < k_blocks) ? p.kchunk : max(k_blocks, 1);
  const int n_chunks = max(1, (k_blocks + kchunk - 1) / kchunk);
  constexpr int n_acc = kExact ? 1 : ACC_STAGES;   // exact: both TMEM buffers belong to one accumulator stage
  const int unit0 = blockIdx.x / 2, unit_stride = gridDim.x / 2;
  // Wave-quantisation fix: the units of the last, partially filled round are split into two 256 x 128 HALF units when
  // that fills the idle CTA pairs (e.g. 4096^2: 256 units on 74 pairs = 3 full rounds + 34 -> 68 half units, 3.5 rounds
  // instead of 4).  The work list is: full units [0, full_units), then half units; all roles walk the same sequence.
  // (Measured: a half unit costs ~0.9 of a full one because it re-reads the whole A tile for half the flops, so the gain
  // is 1-7 %; splitting K instead, with the two partial tiles combined through global memory, measured slower.)
  const int rem = num_units % unit_stride;
  const bool split_tail = p.split_tail && rem > 0 && 2 * rem <= unit_stride && num_units > unit_stride;
  const int full_units = split_tail ? num_units - rem : num_units;
  const int seq_len = full_units + (split_tail ? 2 * rem : 0);

  // work item `seq` -> tile row, first column, half-unit flag
#define PTK_DECODE_UNIT(seq, tm, ncol0, half)                                     \
  int tm, ncol0;                                                                  \
  bool half;                                                                      \
  {                                                                               \
    half = (seq) >= full_units;                                                   \
    const int h_ = ((seq) - full_units) & 1;                                      \
    const int u_ = half ? full_units + ((seq) - full_units) / 2 : (seq);          \
    tm = u_ % m_tiles;                                                            \
    ncol0 = (u_ / m_tiles) * BLOCK_N + (half ? h_ * (BLOCK_N / 2) : 0);           \
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_bh);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of B, completion counted on the LEADER's full barrier =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int sq = unit0; sq < seq_len; sq += unit_stride) {
        PTK_DECODE_UNIT(sq, tm, ncol0, half)
        for (int kb = 0; kb < k_blocks; ++kb) {
          for (int t = 6 - p.terms; t < 6; ++t) {   // one ring stage per piece product (a single pass when terms == 1)
            const int arow = kPieceA[t] * p.a_rows, brow = kPieceB[t] * p.b_rows;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[stage], half ? 2 * (P_A_BYTES + P_B_BYTES / 2) : 2 * P_STAGE_BYTES);
            tma_load_2d_2sm(&tmap_a, &full_bar[stage], smem_a + stage * P_A_BYTES, kb * BLOCK_K,
                            arow + tm * 2 * BLOCK_M + (int)crank * BLOCK_M);
            if (half)  // this CTA's 64 of the 128 B rows of a half unit
              tma_load_2d_2sm(&tmap_bh, &full_bar[stage], smem_b + stage * P_B_BYTES, kb * BLOCK_K,
                              brow + ncol0 + (int)crank * (BLOCK_N / 4));
            else
              tma_load_2d_2sm(&tmap_b, &full_bar[stage], smem_b + stage * P_B_BYTES, kb * BLOCK_K,
                              brow + ncol0 + (int)crank * (BLOCK_N / 2));
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE thread of the leader CTA drives both SMs' tensor cores =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc_full = make_idesc(2 * BLOCK_M, BLOCK_N);
      constexpr uint32_t idesc_half = make_idesc(2 * BLOCK_M, BLOCK_N / 2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr bool exact = kExact;
      const int terms_m1 = p.terms - 1;
      for (int sq = unit0; sq < seq_len; sq += unit_stride) {
        const uint32_t idesc = (sq >= full_units) ? idesc_half : idesc_full;
        for (int ch = 0; ch < n_chunks; ++ch) {   // one TMEM accumulator per accumulation chunk (EpiParams::kchunk)
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * BLOCK_N;
          const uint32_t d_corr = tmem_base + (uint32_t)BLOCK_N;   // exact_main: the correction products' accumulator
          const int kb_n = min(kchunk, k_blocks - ch * kchunk);
          const int n_stages = kb_n * p.terms;  // piece products of one k-block accumulate into the same tile
          int term = 0;                         // position inside the k-block's term sequence (no division in this loop:
          for (int it = 0; it < n_stages; ++it) {   // one thread feeds both SMs' tensor cores, its latency is the pipe's)
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + stage * P_A_BYTES);
            const uint32_t b_addr = smem_u32(smem_b + stage * P_B_BYTES);
            // term order within a k-block: the correction products first, A1 x B1 last (kPieceA / kPieceB)
            const bool to_corr = exact && term != terms_m1;
            const uint32_t d_use = to_corr ? d_corr : d_tmem;
            const bool first = exact ? (to_corr ? it == 0 : it == terms_m1) : it == 0;
            term = (term == terms_m1) ? 0 : term + 1;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_f16_2sm(d_use, make_smem_desc(a_addr + k * UMMA_K * 2), make_smem_desc(b_addr + k * UMMA_K * 2), idesc,
                           (!first || k > 0) ? 1u : 0u);
            }
            umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);  // the stage is free again in BOTH CTAs
            if (++stage == P_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit_2sm_mc(&tmem_full[acc], (uint16_t)0x3);      // accumulator complete in both CTAs' TMEM
          if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the 256-row tile =====
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int sq = unit0; sq < seq_len; sq += unit_stride) {
      PTK_DECODE_UNIT(sq, tm, ncol0, half)
      const int ncols = half ? BLOCK_N / 2 : BLOCK_N;
      const long long row = (long long)tm * 2 * BLOCK_M + (long long)crank * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      float* crow = p.C + row * p.sc0;
      const bool vec_ok = p.sc1 == 1 && (p.sc0 & 3) == 0 && (((uintptr_t)p.C) & 15) == 0 &&
                          (p.Cbf == nullptr || ((p.ldcbf & 3) == 0 && (((uintptr_t)p.Cbf) & 7) == 0));
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch) {
        // chunk 0 combines with the caller's C (beta), later chunks add onto the partial sum this thread stored before;
        // bias / activation / the bf16 copy belong to the completed sum
        const float beta = ch == 0 ? p.beta : 1.0f;
        const bool last = ch == n_chunks - 1;
        const float* bias = last ? p.bias : nullptr;
        const int act = last ? p.act : 0;
        __nv_bfloat16* cbf = last ? p.Cbf : nullptr;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N + c0);
          tmem_ld_32x32b_x32(taddr, r);
          if constexpr (kExact) {
            uint32_t r2[32];
            tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BLOCK_N + c0), r2);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
          } else {
            tmem_ld_wait();
          }
          const long long col0 = (long long)ncol0 + c0;
          // Row-contiguous fast path (warp-uniform): tcgen05.ld hands every lane one ROW of the 32 x 32 block; going through
          // a shared-memory tile turns that into 128 contiguous bytes of one row per quarter warp, so every global
          // load / store instruction of the warp covers 4 full 128-byte lines (the read-modify-write of the accumulation
          // chunks and the bf16 / three-piece operand copies move whole sectors instead of 16-byte shards of 32 lines).
          if (vec_ok && col0 + 32 <= p.N && beta == 0.0f && (cbf == nullptr || p.out_pieces == 1)) {
            // Store-only epilogue (nothing to read back, at most a bf16 copy): each lane keeps its ROW of the block and
            // writes it as eight independent 16-byte groups — no shared-memory round trip, no warp barriers, eight-way
            // instruction-level parallelism through the tanh chains.  The strided 16-byte stores are fire-and-forget; with
            // a bias + tanh + bf16-copy epilogue this path keeps up with the next tile's MMAs where the row-contiguous one
            // below (which pays off as soon as old values must be LOADED) does not (cfg3: 0.43 vs 0.47 ms).
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 v;
                float* vv = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float x = p.alpha * __uint_as_float(r[j + e]);
                  if (bias) x += bias[col0 + j + e];
                  if (act == 1) x = tanhf(x);
                  vv[e] = x;
                }
                *reinterpret_cast<float4*>(crow + col0 + j) = v;
                if (cbf) {
                  __nv_bfloat162 lo = __floats2bfloat162_rn(vv[0], vv[1]), hi = __floats2bfloat162_rn(vv[2], vv[3]);
                  uint2 pk;
                  pk.x = *reinterpret_cast<uint32_t*>(&lo);
                  pk.y = *reinterpret_cast<uint32_t*>(&hi);
                  *reinterpret_cast<uint2*>(cbf + row * p.ldcbf + col0 + j) = pk;
                }
              }
            }
          } else if (vec_ok && col0 + 32 <= p.N) {
            float* tile = epi_tiles + (warp - 4) * (32 * EPI_PITCH);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(tile + lane * EPI_PITCH + j) =
                  make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            __syncwarp();
            const int sub = lane >> 3, cq = (lane & 7) * 4;
            const long long colv = col0 + cq;
            float bb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            if (bias) {
#pragma unroll
              for (int e = 0; e < 4; ++e) bb[e] = bias[colv + e];
            }
            const long long row_base = (long long)tm * 2 * BLOCK_M + (long long)crank * BLOCK_M + q * 32;
            // the old values of all 8 row groups are requested before any of them is used (and before any store of this
            // block, which the compiler could not prove disjoint): one memory round trip per 32 x 32 block, not eight
            float4 old[8];
            if (beta != 0.0f) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const long long grow = row_base + 4 * g + sub;
                old[g] = grow < p.M ? __ldcg(reinterpret_cast<const float4*>(p.C + grow * p.sc0 + colv)) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int rr = 4 * g;
              const long long grow = row_base + rr + sub;
              if (grow < p.M) {
                float4 v = *reinterpret_cast<const float4*>(tile + (rr + sub) * EPI_PITCH + cq);
                float* dst = p.C + grow * p.sc0 + colv;
                float vv[4] = {p.alpha * v.x, p.alpha * v.y, p.alpha * v.z, p.alpha * v.w};
                if (beta != 0.0f) {
                  vv[0] += beta * old[g].x; vv[1] += beta * old[g].y; vv[2] += beta * old[g].z; vv[3] += beta * old[g].w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  vv[e] += bb[e];
                  if (act == 1) vv[e] = tanhf(vv[e]);
                }
                *reinterpret_cast<float4*>(dst) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                if (cbf) {
                  for (int pc = 0; pc < p.out_pieces; ++pc) {   // piece pc = bf16 of what the earlier pieces left over
                    if (pc == 0 && p.out_exp != PTK_NO_EXP) {     // aligned leading piece (exact in bf16)
#pragma unroll
                      for (int e = 0; e < 4; ++e) {
                        const float lead = rintf(vv[e] * p.out_scale) * p.out_inv;
                        vv[e] -= lead;
                        reinterpret_cast<float*>(&v)[e] = lead;
                      }
                      __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                      uint2 pk;
                      pk.x = *reinterpret_cast<uint32_t*>(&lo);
                      pk.y = *reinterpret_cast<uint32_t*>(&hi);
                      *reinterpret_cast<uint2*>(cbf + ((long long)pc * p.cbf_rows + grow) * p.ldcbf + colv) = pk;
                      continue;
                    }
                    __nv_bfloat162 lo = __floats2bfloat162_rn(vv[0], vv[1]), hi = __floats2bfloat162_rn(vv[2], vv[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(cbf + ((long long)pc * p.cbf_rows + grow) * p.ldcbf + colv) = pk;
                    vv[0] -= __bfloat162float(lo.x); vv[1] -= __bfloat162float(lo.y);
                    vv[2] -= __bfloat162float(hi.x); vv[3] -= __bfloat162float(hi.y);
                  }
                }
              }
            }
            __syncwarp();
          } else if (row_ok && col0 < p.N) {
            {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const long long col = col0 + j;
                if (col < p.N) {
                  float* dst = crow + col * p.sc1;
                  float x = p.alpha * __uint_as_float(r[j]);
                  if (beta != 0.0f) x += beta * (*dst);
                  if (bias) x += bias[col];
                  if (act == 1) x = tanhf(x);
                  *dst = x;
                  if (cbf) {
                    for (int pc = 0; pc < p.out_pieces; ++pc) {
                      const __nv_bfloat16 b = (pc == 0 && p.out_exp != PTK_NO_EXP)
                                                  ? __float2bfloat16_rn(rintf(x * p.out_scale) * p.out_inv)
                                                  : __float2bfloat16_rn(x);
                      cbf[((long long)pc * p.cbf_rows + row) * p.ldcbf + col] = b;
                      x -= __bfloat162float(b);
                    }
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);  // 8 arrivals (4 warps x 2 CTAs) free the accumulator
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
    }
  }
#undef PTK_DECODE_UNIT
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

// ---- fp32 -> bf16 operand staging: dst[r][c] (row-major, pitch ld) = bf16(src[r*sr + c*sc]) ----------------------------
// Tiled through shared memory so that both the read (along whichever source stride is 1) and the write (along c) coalesce.
__global__ void __launch_bounds__(256) convert_bf16_kernel(const float* __restrict__ src, long long sr, long long sc,
                                                           __nv_bfloat16* __restrict__ dst, long long ld, long long R,
                                                           long long Cc) {
  // 64 x 64 tile: reads walk the source's unit-stride direction (128 B per warp), writes are packed bf16x2 along c
  // (128 B per warp-row).  ld is even and dst 4-byte aligned (workspace rows are padded to 8 elements).
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool col_fast = (sc == 1) || (sr != 1);
  if (col_fast) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + ty + 8 * i, c = c0 + tx + 32 * h;
        tile[ty + 8 * i][tx + 32 * h] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  } else {  // source is row-fast (sr == 1): read with threads along r, transpose through shared memory
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + tx + 32 * h, c = c0 + ty + 8 * i;
        tile[tx + 32 * h][ty + 8 * i] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    if (r < R && c < Cc) {
      const float lo = tile[ty + 8 * i][2 * tx], hi = tile[ty + 8 * i][2 * tx + 1];
      if (c + 1 < Cc || c + 1 < ld) {
        *reinterpret_cast<__nv_bfloat162*>(dst + r * ld + c) = __floats2bfloat162_rn(lo, (c + 1 < Cc) ? hi : 0.0f);
      } else {
        dst[r * ld + c] = __float2bfloat16_rn(lo);
      }
    }
  }
}

// ---- fp32 -> 3 x bf16 operand split: piece i of src[r*sr + c*sc] goes to dst[(i * piece_rows + r) * ld + c] ----------------
// x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2): the residuals are exact in fp32, so x1 + x2 + x3 carries 24
// mantissa bits of x.  Same 64 x 64 shared-memory tile as convert_bf16_kernel (coalesced reads along either source stride).
__global__ void __launch_bounds__(256) split_bf16x3_kernel(const float* __restrict__ src, long long sr, long long sc,
                                                           __nv_bfloat16* __restrict__ dst, long long ld, long long R,
                                                           long long Cc, long long piece_rows) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool col_fast = (sc == 1) || (sr != 1);
  if (col_fast) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + ty + 8 * i, c = c0 + tx + 32 * h;
        tile[ty + 8 * i][tx + 32 * h] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + tx + 32 * h, c = c0 + ty + 8 * i;
        tile[tx + 32 * h][ty + 8 * i] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    if (r < R && c < Cc) {
      float lo = tile[ty + 8 * i][2 * tx], hi = (c + 1 < Cc) ? tile[ty + 8 * i][2 * tx + 1] : 0.0f;
#pragma unroll
      for (int pc = 0; pc < 3; ++pc) {
        const __nv_bfloat16 bl = __float2bfloat16_rn(lo), bh = __float2bfloat16_rn(hi);
        __nv_bfloat16* d = dst + (pc * piece_rows + r) * ld + c;
        if (c + 1 < ld) {
          __nv_bfloat162 pk;
          pk.x = bl;
          pk.y = bh;
          *reinterpret_cast<__nv_bfloat162*>(d) = pk;
        } else {
          *d = bl;
        }
        lo -= __bfloat162float(bl);
        hi -= __bfloat162float(bh);
      }
    }
  }
}

// ---- error-free leading piece (fp32-accurate mode, "exact main term") ---------------------------------------------------
// The tensor core truncates when it adds into its fp32 accumulator; over a chain of MMAs that is a systematic shrink of the
// result (~1e-7 per MMA of the chain, measured) which, unlike rounding noise, adds up coherently through chained layers.
// Truncation cannot bite when every partial sum is exactly representable: the LEADING piece of each operand is therefore
// taken as an integer multiple of a per-row power of two, x1 = rint(x * 2^s) * 2^-s with |rint| <= 2^b (s = b - 1 - ilogb of
// the row's largest magnitude; a row of A, a column of B).  All products A1[i,k] * B1[k,j] of one output element are then
// integers (<= 2^2b, b = 7: lead_bits_for) on the common unit 2^-(s_i + s_j), and they sum exactly in fp32 as long as the
// partial sums stay below 2^24 units; the operand carries b + 16 = 23 bits plus sign through its three pieces.  The remainder x - x1 is
// exact in fp32 and is split into two ordinary bf16 pieces; the five correction products go to a second accumulator, whose
// own truncation shrink is 2^-7 of the total.
// Pass 1: largest magnitude of every row, as the bit pattern of a non-negative float (ordered like an unsigned integer):
// 64 x 64 tiles through shared memory exactly like the split kernels (coalesced whichever source stride is 1), four
// threads per tile row, one atomicMax per tile row.  `maxbits` must be zeroed beforehand.  NaN / inf count as "no scale".
__global__ void __launch_bounds__(256) row_absmax_kernel(const float* __restrict__ src, long long sr, long long sc,
                                                         long long R, long long Cc, unsigned int* __restrict__ maxbits) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool col_fast = (sc == 1) || (sr != 1);
  if (col_fast) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + ty + 8 * i, c = c0 + tx + 32 * h;
        tile[ty + 8 * i][tx + 32 * h] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + tx + 32 * h, c = c0 + ty + 8 * i;
        tile[tx + 32 * h][ty + 8 * i] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  }
  __syncthreads();
  const int row = threadIdx.x >> 2, part = threadIdx.x & 3;
  float m = 0.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float v = fabsf(tile[row][part * 16 + j]);
    m = (v > m) ? v : m;   // NaN never wins; +inf does (and then disables the scale below)
  }
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  if (part == 0 && r0 + row < R && m > 0.0f) atomicMax(maxbits + r0 + row, __float_as_uint(m));
}

// per-row scale exponent from the row maximum: (lead_bits - 1) - ilogb(max), so that |x| * 2^s < 2^lead_bits; 0 for empty /
// non-finite rows
__device__ __forceinline__ int scale_exp_of(unsigned int bits, int lead_bits) {
  const float m = __uint_as_float(bits);
  return (m > 0.0f && m < __int_as_float(0x7f800000)) ? (lead_bits - 1) - ilogbf(m) : 0;
}

// piece 0 = rint(x * 2^s[r]) * 2^-s[r] (exact in bf16), pieces 1, 2 = bf16 split of the exact remainder; same tiling and
// output layout as split_bf16x3_kernel.  fixed_exp != INT_MIN: use that exponent for every row (operands known to lie in
// [-1, 1], e.g. tanh outputs written by a previous product's epilogue) instead of sexp.
__global__ void __launch_bounds__(256) split_aligned_kernel(const float* __restrict__ src, long long sr, long long sc,
                                                            __nv_bfloat16* __restrict__ dst, long long ld, long long R,
                                                            long long Cc, long long piece_rows,
                                                            const unsigned int* __restrict__ maxbits, int lead_bits) {
  __shared__ float tile[64][65];
  const long long r0 = (long long)blockIdx.y * 64, c0 = (long long)blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const bool col_fast = (sc == 1) || (sr != 1);
  if (col_fast) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + ty + 8 * i, c = c0 + tx + 32 * h;
        tile[ty + 8 * i][tx + 32 * h] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long r = r0 + tx + 32 * h, c = c0 + ty + 8 * i;
        tile[tx + 32 * h][ty + 8 * i] = (r < R && c < Cc) ? src[r * sr + c * sc] : 0.0f;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + ty + 8 * i, c = c0 + 2 * tx;
    if (r < R && c < Cc) {
      int sx = scale_exp_of(maxbits[r], lead_bits);
      sx = sx > 126 ? 126 : (sx < -126 ? -126 : sx);   // 2^sx and 2^-sx must be normal floats (rows of denormal size lose bits)
      const float up = __int_as_float((127 + sx) << 23), dn = __int_as_float((127 - sx) << 23);
      float v[2] = {tile[ty + 8 * i][2 * tx], (c + 1 < Cc) ? tile[ty + 8 * i][2 * tx + 1] : 0.0f};
      __nv_bfloat16 pc[3][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float lead = rintf(v[e] * up) * dn;   // |rint| <= 128: exact in bf16; x - lead exact in fp32
        pc[0][e] = __float2bfloat16_rn(lead);
        float rem = v[e] - lead;
        if (!(fabsf(v[e]) < __int_as_float(0x7f800000))) rem = 0.0f;  // inf / NaN ride in the leading piece only
        pc[1][e] = __float2bfloat16_rn(rem);
        rem -= __bfloat162float(pc[1][e]);
        pc[2][e] = __float2bfloat16_rn(rem);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        __nv_bfloat16* d = dst + (k * piece_rows + r) * ld + c;
        if (c + 1 < ld) {
          __nv_bfloat162 pk;
          pk.x = pc[k][0];
          pk.y = pc[k][1];
          *reinterpret_cast<__nv_bfloat162*>(d) = pk;
        } else {
          *d = pc[k][0];
        }
      }
    }
  }
}

ptk_status make_tmap(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ptk::drv().TensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr,
                                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return ptk::check_cu(r, "cuTensorMapEncodeTiled");
}

inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }

}  // namespace

namespace ptk {

size_t gemm_tc_workspace(int64_t M, int64_t N, int64_t K) {
  long long Kp = round_up(K, 8);
  return (size_t)(round_up(M * Kp * 2, 256) + round_up(N * Kp * 2, 256) + 256);
}

size_t gemm_tc_split_workspace(int64_t M, int64_t N, int64_t K) {
  long long Kp = round_up(K, 8), Mp = round_up(M, 256), Np = round_up(N, 256);
  return (size_t)(round_up(3 * Mp * Kp * 2, 256) + round_up(3 * Np * Kp * 2, 256) + round_up(4 * (M + N), 256) + 256);
}

// error-free leading pieces on/off for the fp32-accurate mode (PTK_GEMM_EXACT=0: plain bf16x3 split, one accumulator)
// Leading-piece width: |rint(x * 2^s)| <= 2^7.  Products are then integers <= 2^14 and every partial sum below 2^24 units is
// exact: always for K <= 1024, and for longer contractions unless more than a thousand near-maximal products line up in
// sign — past 2^24 the accumulator merely drops its lowest one or two unit bits (a 2^-24-level effect, no worse than the
// plain split, and only on such data).  Measured alternative (width shrinking with K so that exactness is unconditional:
// 6 bits at K = 4096, 5 at 8192): max error 2.5e-6 / 9.3e-6 of the output scale instead of 6e-7 — the operand then only
// carries b + 16 bits — so the width stays 7 and the accumulation is cut into chunks only beyond K = 16384.
static int lead_bits_for(int64_t K) {
  (void)K;
  return 7;
}

static int exact_main_default() {
  static int g = -1;
  if (g < 0) {
    const char* e = getenv("PTK_GEMM_EXACT");
    g = (e && e[0] == '0') ? 0 : 1;
  }
  return g;
}

static int g_cluster = -1;  // -1: read PTK_GEMM_CLUSTER once (default 2)

// A_f32 (any strides) OR A_bf16 (row-major, pitch lda_bf16 elements, multiple of 8, 16-byte aligned base) must be given.
ptk_status gemm_tc_ex(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sa0, int64_t sa1,
                      const void* A_bf16, int64_t lda_bf16, const float* B, int64_t sb0, int64_t sb1, float beta, float* C,
                      int64_t sc0, int64_t sc1, const float* bias, int act, void* C_bf16, int64_t ldc_bf16, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
  if (M == 0 || N == 0) return PTK_OK;
  if (M > 2147483647LL || N > 2147483647LL || K > 2147483647LL) return fail(PTK_ERR_ARG, "gemm_tc: dims exceed int32");
  if (workspace == nullptr || workspace_bytes < gemm_tc_workspace(M, N, K))
    return fail(PTK_ERR_ARG, "gemm_tc: workspace too small (see ptk_gemm_workspace_bytes)");
  if (g_cluster < 0) {
    const char* e = getenv("PTK_GEMM_MODE");  // 1 = single CTA, 2 = CTA pair sharing B by TMA multicast, 3 = cta_group::2 UMMA
    g_cluster = (e && e[0] >= '1' && e[0] <= '3') ? (e[0] - '0') : 3;
  }
  const long long Kp = round_up(K, 8);
  uintptr_t w = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  __nv_bfloat16* Abf = reinterpret_cast<__nv_bfloat16*>(w);
  __nv_bfloat16* Bbf = reinterpret_cast<__nv_bfloat16*>(w + round_up(M * Kp * 2, 256));
  long long lda = Kp;
  if (A_bf16 != nullptr) {
    if (lda_bf16 % 8 != 0 || ((uintptr_t)A_bf16 & 15) != 0) return fail(PTK_ERR_ARG, "gemm_tc: misaligned bf16 A operand");
    Abf = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(A_bf16));
    lda = lda_bf16;
  } else {
    dim3 ga((unsigned)((K + 63) / 64), (unsigned)((M + 63) / 64));
    convert_bf16_kernel<<<ga, 256, 0, st>>>(A, sa0, sa1, Abf, Kp, M, K);
  }
  {
    // B[K,N] -> Bt[N,K]: dst row index = n (source stride sb1), dst col index = k (source stride sb0)
    dim3 gb((unsigned)((K + 63) / 64), (unsigned)((N + 63) / 64));
    convert_bf16_kernel<<<gb, 256, 0, st>>>(B, sb1, sb0, Bbf, Kp, N, K);
    PTK_LAUNCH_CHECK("convert_bf16");
  }
  const int cluster = g_cluster;
  CUtensorMap ta, tb, tbh;
  ptk_status s;
  if ((s = make_tmap(&ta, Abf, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BLOCK_M)) != PTK_OK) return s;
  if ((s = make_tmap(&tbh, Bbf, (uint64_t)N, (uint64_t)K, (uint64_t)Kp, BLOCK_N / 4)) != PTK_OK) return s;
  if ((s = make_tmap(&tb, Bbf, (uint64_t)N, (uint64_t)K, (uint64_t)Kp, cluster >= 2 ? BLOCK_N / 2 : BLOCK_N)) != PTK_OK) return s;
  EpiParams p;
  p.alpha = alpha; p.beta = beta; p.C = C; p.sc0 = sc0; p.sc1 = sc1; p.bias = bias; p.act = act;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.Cbf = reinterpret_cast<__nv_bfloat16*>(C_bf16);
  p.ldcbf = ldc_bf16;
  static int g_split = -1;
  if (g_split < 0) {
    const char* e = getenv("PTK_GEMM_SPLIT");
    g_split = (e && e[0] == '0') ? 0 : 1;
  }
  p.split_tail = g_split;
  p.terms = 1; p.a_rows = 0; p.b_rows = 0; p.kchunk = 0; p.out_pieces = 1; p.cbf_rows = 0; p.exact_main = 0; p.out_exp = PTK_NO_EXP; p.out_scale = 1.0f; p.out_inv = 1.0f;
  static bool attr_set = false;
  if (!attr_set) {
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M), n_tiles = (int)((N + BLOCK_N - 1) / BLOCK_N);
  const int sms = std::max(2, ptk::sm_count());
  if (cluster >= 2) {
    const int units = ((m_tiles + 1) / 2) * n_tiles;
    const int grid = 2 * std::max(1, std::min(units, sms / 2));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = cluster == 3 ? P_SMEM_BYTES : SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cluster == 3) PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_pair_kernel<false>, ta, tb, tbh, p));
    else PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<2>, ta, tb, p));
  } else {
    const int grid = std::max(1, std::min(m_tiles * n_tiles, sms));
    gemm_bf16_tc_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ta, tb, p);
  }
  PTK_LAUNCH_CHECK("gemm_bf16_tc");
  return PTK_OK;
}

// accumulation chunk of the fp32-accurate modes: 8 k-blocks (K = 512) keeps the tensor core's truncation bias near 1e-6 of
// the output scale; PTK_GEMM_KCHUNK=<k-blocks> overrides (0 = one chunk)
static int split_kchunk(int64_t K, int exact) {
  static int g_kchunk = -1;
  if (g_kchunk < 0) {
    const char* e = getenv("PTK_GEMM_KCHUNK");
    g_kchunk = e ? atoi(e) : -1;
  }
  const long long kb = (K + BLOCK_K - 1) / BLOCK_K;
  if (exact) {
    const long long cap = 256;   // k-blocks (K = 16384) per accumulation, see lead_bits_for
    const long long c = g_kchunk > 0 ? std::min<long long>(g_kchunk, cap) : cap;
    return kb > c ? (int)c : 0;
  }
  const int c = g_kchunk >= 0 ? g_kchunk : 8;
  return (c > 0 && kb > c + c / 4) ? c : 0;
}

// Operand staging on its own (so that an operand that does not change between calls is staged ONCE): dst = `pieces` (1 | 3)
// bf16 matrices [R, Cc] stacked with a pitch of piece_rows rows, row pitch ld elements, from fp32 src[r*sr + c*sc].
ptk_status stage_operand(const float* src, int64_t sr, int64_t sc, int64_t R, int64_t Cc, int pieces, void* dst, int64_t ld,
                         int64_t piece_rows, int aligned, int* sexp, cudaStream_t st) {
  if (pieces != 1 && pieces != 3) return fail(PTK_ERR_ARG, "stage_operand: pieces must be 1 or 3");
  if (aligned && (pieces != 3 || sexp == nullptr)) return fail(PTK_ERR_ARG, "stage_operand: aligned staging needs 3 pieces and the exponent scratch");
  if (ld % 8 != 0 || ld < Cc || ((uintptr_t)dst & 15) != 0) return fail(PTK_ERR_ARG, "stage_operand: pitch must be a multiple of 8 >= cols, base 16-byte aligned");
  if (R == 0 || Cc == 0) return PTK_OK;
  dim3 g((unsigned)((Cc + 63) / 64), (unsigned)((R + 63) / 64));
  if (pieces == 1) {
    convert_bf16_kernel<<<g, 256, 0, st>>>(src, sr, sc, (__nv_bfloat16*)dst, ld, R, Cc);
  } else if (aligned) {
    PTK_CUDA(cudaMemsetAsync(sexp, 0, (size_t)R * 4, st));
    row_absmax_kernel<<<g, 256, 0, st>>>(src, sr, sc, R, Cc, reinterpret_cast<unsigned int*>(sexp));
    split_aligned_kernel<<<g, 256, 0, st>>>(src, sr, sc, (__nv_bfloat16*)dst, ld, R, Cc, piece_rows,
                                            reinterpret_cast<const unsigned int*>(sexp), lead_bits_for(Cc));
  } else {
    split_bf16x3_kernel<<<g, 256, 0, st>>>(src, sr, sc, (__nv_bfloat16*)dst, ld, R, Cc, piece_rows);
  }
  PTK_LAUNCH_CHECK("stage_operand");
  return PTK_OK;
}

// The cta_group::2 kernel over operands that are ALREADY staged (see stage_operand): A_stage = pieces of A [M,K] (K-major,
// pitch lda, piece pitch a_rows rows), B_stage = pieces of B^T [N,K]; terms 1 (plain bf16) | 3 | 6.  C_stage (optional)
// receives `out_pieces` (1 | 3) staged pieces of the RESULT [M,N] (pitch ldc_stage, piece pitch c_rows) — the A operand
// of the next product of a chain / recurrence, so that only the very first operand is ever staged by a separate pass.
ptk_status gemm_tc_staged(int64_t M, int64_t N, int64_t K, float alpha, const void* A_stage, int64_t lda, int64_t a_rows,
                          const void* B_stage, int64_t ldb, int64_t b_rows, int terms, float beta, float* C, int64_t sc0,
                          int64_t sc1, const float* bias, int act, void* C_stage, int64_t ldc_stage, int64_t c_rows,
                          int out_pieces, int exact_main, int out_exp, cudaStream_t st) {
  if (M == 0 || N == 0) return PTK_OK;
  if (terms != 1 && terms != 3 && terms != 6) return fail(PTK_ERR_ARG, "gemm_tc_staged: terms must be 1, 3 or 6");
  if (M > 500000000LL || N > 500000000LL || K > 2147483647LL || K <= 0) return fail(PTK_ERR_ARG, "gemm_tc_staged: bad dims");
  if (lda % 8 || ldb % 8 || ((uintptr_t)A_stage & 15) || ((uintptr_t)B_stage & 15))
    return fail(PTK_ERR_ARG, "gemm_tc_staged: operand pitch must be a multiple of 8 elements, base 16-byte aligned");
  if (C_stage != nullptr && (out_pieces != 1 && out_pieces != 3)) return fail(PTK_ERR_ARG, "gemm_tc_staged: out_pieces must be 1 or 3");
  if (C_stage != nullptr && (ldc_stage % 8 || ((uintptr_t)C_stage & 15))) return fail(PTK_ERR_ARG, "gemm_tc_staged: misaligned C_stage");
  const int pa = terms == 1 ? 1 : 3;
  CUtensorMap ta, tb, tbh;
  ptk_status s;
  const uint64_t a_total = (uint64_t)((pa - 1) * a_rows + M), b_total = (uint64_t)((pa - 1) * b_rows + N);
  if ((s = make_tmap(&ta, A_stage, a_total, (uint64_t)K, (uint64_t)lda, BLOCK_M)) != PTK_OK) return s;
  if ((s = make_tmap(&tbh, B_stage, b_total, (uint64_t)K, (uint64_t)ldb, BLOCK_N / 4)) != PTK_OK) return s;
  if ((s = make_tmap(&tb, B_stage, b_total, (uint64_t)K, (uint64_t)ldb, BLOCK_N / 2)) != PTK_OK) return s;
  EpiParams p;
  p.alpha = alpha; p.beta = beta; p.C = C; p.sc0 = sc0; p.sc1 = sc1; p.bias = bias; p.act = act;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.Cbf = reinterpret_cast<__nv_bfloat16*>(C_stage);
  p.ldcbf = ldc_stage;
  p.out_pieces = C_stage ? out_pieces : 1;
  p.cbf_rows = c_rows;
  p.exact_main = 0; p.out_exp = PTK_NO_EXP; p.out_scale = 1.0f; p.out_inv = 1.0f;
  p.split_tail = 1;
  p.terms = terms; p.a_rows = terms == 1 ? 0 : (int)a_rows; p.b_rows = terms == 1 ? 0 : (int)b_rows;
  p.exact_main = (terms != 1 && exact_main) ? 1 : 0;
  p.out_exp = (C_stage && out_pieces == 3) ? out_exp : PTK_NO_EXP;
  if (p.out_exp != PTK_NO_EXP) {
    if (p.out_exp < -100 || p.out_exp > 100) return fail(PTK_ERR_ARG, "gemm_tc_staged: out_exp out of range");
    p.out_scale = ldexpf(1.0f, p.out_exp);
    p.out_inv = ldexpf(1.0f, -p.out_exp);
  }
  p.kchunk = terms == 1 ? 0 : split_kchunk(K, p.exact_main);
  static bool attr_set = false;
  if (!attr_set) {
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M), n_tiles = (int)((N + BLOCK_N - 1) / BLOCK_N);
  const int sms = std::max(2, ptk::sm_count());
  const int units = ((m_tiles + 1) / 2) * n_tiles;
  const int grid = 2 * std::max(1, std::min(units, sms / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.exact_main) PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_pair_kernel<true>, ta, tb, tbh, p));
  else PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_pair_kernel<false>, ta, tb, tbh, p));
  PTK_LAUNCH_CHECK("gemm_tc_staged");
  return PTK_OK;
}

// fp32-accurate product on the tensor cores: both operands split into three bf16 pieces, `terms` (3 or 6) piece products
// per k-block accumulated in fp32 TMEM by the cta_group::2 kernel (see EpiParams::terms).
ptk_status gemm_tc_split(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sa0, int64_t sa1, const float* B,
                         int64_t sb0, int64_t sb1, float beta, float* C, int64_t sc0, int64_t sc1, const float* bias, int act,
                         int terms, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (M == 0 || N == 0) return PTK_OK;
  if (terms != 3 && terms != 6) return fail(PTK_ERR_ARG, "gemm_tc_split: terms must be 3 or 6");
  if (M > 500000000LL || N > 500000000LL || K > 2147483647LL) return fail(PTK_ERR_ARG, "gemm_tc_split: dims exceed int32");
  if (workspace == nullptr || workspace_bytes < gemm_tc_split_workspace(M, N, K))
    return fail(PTK_ERR_ARG, "gemm_tc_split: workspace too small (see ptk_gemm_split_workspace_bytes)");
  const long long Kp = round_up(K, 8), Mp = round_up(M, 256), Np = round_up(N, 256);
  uintptr_t w = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  __nv_bfloat16* Abf = reinterpret_cast<__nv_bfloat16*>(w);
  __nv_bfloat16* Bbf = reinterpret_cast<__nv_bfloat16*>(w + round_up(3 * Mp * Kp * 2, 256));
  const int exact = terms == 6 ? exact_main_default() : 0;   // (3 terms need the 8-bit leading pieces of the plain split)
  {
    int* sexp = reinterpret_cast<int*>(w + round_up(3 * Mp * Kp * 2, 256) + round_up(3 * Np * Kp * 2, 256));
    ptk_status ss;
    if ((ss = stage_operand(A, sa0, sa1, M, K, 3, Abf, Kp, Mp, exact, sexp, st)) != PTK_OK) return ss;
    if ((ss = stage_operand(B, sb1, sb0, N, K, 3, Bbf, Kp, Np, exact, sexp + M, st)) != PTK_OK) return ss;  // B[K,N] -> Bt[N,K]
  }
  CUtensorMap ta, tb, tbh;
  ptk_status s;
  // rows past M (N) inside a piece hold whatever the workspace held: they only reach output rows (columns) the epilogue
  // masks, never a stored element; columns past K are zero-filled by TMA (the map's extent is K)
  if ((s = make_tmap(&ta, Abf, (uint64_t)(3 * Mp), (uint64_t)K, (uint64_t)Kp, BLOCK_M)) != PTK_OK) return s;
  if ((s = make_tmap(&tbh, Bbf, (uint64_t)(3 * Np), (uint64_t)K, (uint64_t)Kp, BLOCK_N / 4)) != PTK_OK) return s;
  if ((s = make_tmap(&tb, Bbf, (uint64_t)(3 * Np), (uint64_t)K, (uint64_t)Kp, BLOCK_N / 2)) != PTK_OK) return s;
  EpiParams p;
  p.alpha = alpha; p.beta = beta; p.C = C; p.sc0 = sc0; p.sc1 = sc1; p.bias = bias; p.act = act;
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.Cbf = nullptr;
  p.ldcbf = 0;
  p.out_pieces = 1; p.cbf_rows = 0; p.exact_main = 0; p.out_exp = PTK_NO_EXP; p.out_scale = 1.0f; p.out_inv = 1.0f;
  p.split_tail = 1;
  p.terms = terms; p.a_rows = (int)Mp; p.b_rows = (int)Np;
  p.exact_main = exact;
  p.kchunk = split_kchunk(K, exact);
  static bool attr_set = false;
  if (!attr_set) {
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    PTK_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (int)((M + BLOCK_M - 1) / BLOCK_M), n_tiles = (int)((N + BLOCK_N - 1) / BLOCK_N);
  const int sms = std::max(2, ptk::sm_count());
  const int units = ((m_tiles + 1) / 2) * n_tiles;
  const int grid = 2 * std::max(1, std::min(units, sms / 2));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = P_SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.exact_main) PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_pair_kernel<true>, ta, tb, tbh, p));
  else PTK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_pair_kernel<false>, ta, tb, tbh, p));
  PTK_LAUNCH_CHECK("gemm_bf16x3_tc");
  return PTK_OK;
}

ptk_status gemm_tc(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sa0, int64_t sa1,
                   const float* B, int64_t sb0, int64_t sb1, float beta, float* C, int64_t sc0, int64_t sc1,
                   const float* bias, int act, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  return gemm_tc_ex(M, N, K, alpha, A, sa0, sa1, nullptr, 0, B, sb0, sb1, beta, C, sc0, sc1, bias, act, nullptr, 0, workspace,
                    workspace_bytes, st);
}

}  // namespace ptk

extern "C" ptk_status ptk_gemm_tc_ex(int64_t M, int64_t N, int64_t K, double alpha, const void* A_f32, int64_t sa0,
                                     int64_t sa1, const void* A_bf16, int64_t lda_bf16, const void* B_f32, int64_t sb0,
                                     int64_t sb1, double beta, void* C, int64_t sc0, int64_t sc1, const void* bias, int act,
                                     void* C_bf16, int64_t ldc_bf16, void* workspace, size_t workspace_bytes, void* stream) {
  PTK_REQUIRE_INIT();
  if (A_f32 == nullptr && A_bf16 == nullptr) return ptk::fail(PTK_ERR_ARG, "ptk_gemm_tc_ex: no A operand");
  return ptk::gemm_tc_ex(M, N, K, (float)alpha, (const float*)A_f32, sa0, sa1, A_bf16, lda_bf16, (const float*)B_f32, sb0, sb1,
                         (float)beta, (float*)C, sc0, sc1, (const float*)bias, act, C_bf16, ldc_bf16, workspace,
                         workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t ptk_gemm_split_workspace_bytes(int64_t M, int64_t N, int64_t K) { return ptk::gemm_tc_split_workspace(M, N, K); }

extern "C" ptk_status ptk_gemm_tc_split(int64_t M, int64_t N, int64_t K, double alpha, const void* A_f32, int64_t sa0,
                                        int64_t sa1, const void* B_f32, int64_t sb0, int64_t sb1, double beta, void* C,
                                        int64_t sc0, int64_t sc1, const void* bias, int act, int terms, void* workspace,
                                        size_t workspace_bytes, void* stream) {
  PTK_REQUIRE_INIT();
  if (A_f32 == nullptr || B_f32 == nullptr) return ptk::fail(PTK_ERR_ARG, "ptk_gemm_tc_split: null operand");
  return ptk::gemm_tc_split(M, N, K, (float)alpha, (const float*)A_f32, sa0, sa1, (const float*)B_f32, sb0, sb1, (float)beta,
                            (float*)C, sc0, sc1, (const float*)bias, act, terms, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

extern "C" size_t ptk_stage_bytes(int64_t rows, int64_t cols, int pieces) {
  const long long ld = (cols + 7) / 8 * 8, pr = (rows + 255) / 256 * 256;
  const size_t mats = (size_t)(pieces <= 1 ? rows : 3 * pr) * (size_t)ld * 2;
  return (mats + 255) / 256 * 256 + (pieces <= 1 ? 0 : (size_t)rows * 4) + 256;   // + the per-row exponents of an aligned split
}

extern "C" ptk_status ptk_stage_operand(const void* src_f32, int64_t sr, int64_t sc, int64_t rows, int64_t cols, int pieces,
                                        int aligned, void* dst, int64_t ld, int64_t piece_rows, void* stream) {
  PTK_REQUIRE_INIT();
  if (src_f32 == nullptr || dst == nullptr) return ptk::fail(PTK_ERR_ARG, "ptk_stage_operand: null pointer");
  int* sexp = nullptr;
  if (aligned) {
    if (pieces != 3 || ld != (cols + 7) / 8 * 8 || piece_rows != (rows + 255) / 256 * 256)
      return ptk::fail(PTK_ERR_ARG, "ptk_stage_operand: aligned staging uses the default pitch / piece pitch of ptk_stage_bytes");
    const size_t mats = (size_t)(3 * piece_rows) * (size_t)ld * 2;
    sexp = reinterpret_cast<int*>((char*)dst + (mats + 255) / 256 * 256);
  }
  return ptk::stage_operand((const float*)src_f32, sr, sc, rows, cols, pieces, dst, ld, piece_rows, aligned, sexp,
                            (cudaStream_t)stream);
}

extern "C" ptk_status ptk_gemm_tc_staged(int64_t M, int64_t N, int64_t K, double alpha, const void* A_stage, int64_t lda,
                                         int64_t a_rows, const void* B_stage, int64_t ldb, int64_t b_rows, int terms,
                                         double beta, void* C, int64_t sc0, int64_t sc1, const void* bias, int act,
                                         void* C_stage, int64_t ldc_stage, int64_t c_rows, int out_pieces, int exact_main,
                                         int out_exp, void* stream) {
  PTK_REQUIRE_INIT();
  if (A_stage == nullptr || B_stage == nullptr || C == nullptr) return ptk::fail(PTK_ERR_ARG, "ptk_gemm_tc_staged: null operand");
  return ptk::gemm_tc_staged(M, N, K, (float)alpha, A_stage, lda, a_rows, B_stage, ldb, b_rows, terms, (float)beta, (float*)C,
                             sc0, sc1, (const float*)bias, act, C_stage, ldc_stage, c_rows, out_pieces, exact_main,
                             out_exp == PTK_STAGE_NO_EXP ? PTK_NO_EXP : out_exp, (cudaStream_t)stream);
}

extern "C" int ptk_gemm_exact_main_default(void) { return ptk::exact_main_default(); }

extern "C" int ptk_gemm_lead_bits(int64_t K) { return ptk::lead_bits_for(K); }
