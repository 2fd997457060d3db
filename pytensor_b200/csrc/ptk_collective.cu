// One-shot all-reduce (sum) of a SMALL vector over NVLink peer memory — the exchange step of the batch-sharded logp+grad
// evaluation (one message of 1 + P = 75 floats per evaluation, purely latency-bound).
//
// Every rank owns a symmetric buffer (same layout on all ranks, peer-mapped through NVLink/NVSwitch; the pointers come from
// torch.distributed's symmetric-memory rendezvous, which is only plumbing):
//     data [2 parities][world][nmax] elements, then flags [2 parities][world] uint32.
// One CTA per rank:  (1) PUSH its vector into slot [parity][rank] of EVERY rank's buffer with plain remote stores,
// (2) fence + publish a release flag = epoch in every rank's flag slot, (3) spin (acquire) until all `world` flags of its OWN
// buffer carry this epoch, (4) sum the `world` vectors in rank order (bit-identical result on every rank).  The epoch lives
// in device memory, so the kernel can sit inside a captured CUDA graph.  Two parities make slot reuse safe: a rank can only
// reach epoch e+2 after it saw every peer's flag of epoch e+1, which the peer publishes after it finished reading epoch e.
// The reference has no collective at all; NCCL's all-reduce is the baseline this replaces for tiny messages.
#include "ptk_common.h"

namespace {

constexpr int kMaxWorld = 16;

struct Peers {
  void* buf[kMaxWorld];
};

template <typename T>
__global__ void __launch_bounds__(256) allreduce_oneshot_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n,
                                                                Peers peers, int rank, int world, int64_t nmax,
                                                                unsigned int* epoch_ctr) {
  __shared__ unsigned int s_epoch;
  if (threadIdx.x == 0) s_epoch = *epoch_ctr + 1u;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const int parity = (int)(epoch & 1u);
  const size_t data_elems = (size_t)2 * world * nmax;
  // (1) push
  for (int p = 0; p < world; ++p) {
    T* dst = reinterpret_cast<T*>(peers.buf[p]) + ((size_t)parity * world + rank) * nmax;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  // (2) publish
  if ((int)threadIdx.x < world) {
    unsigned int* flags = reinterpret_cast<unsigned int*>(reinterpret_cast<T*>(peers.buf[threadIdx.x]) + data_elems);
    unsigned int* f = flags + (size_t)parity * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");
  }
  // (3) wait for every rank's contribution in MY buffer
  if ((int)threadIdx.x < world) {
    unsigned int* flags = reinterpret_cast<unsigned int*>(reinterpret_cast<T*>(peers.buf[rank]) + data_elems);
    unsigned int* f = flags + (size_t)parity * world + threadIdx.x;
    unsigned int v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while (v != epoch);
  }
  __syncthreads();
  // (4) reduce in rank order
  const T* mine = reinterpret_cast<const T*>(peers.buf[rank]) + (size_t)parity * world * nmax;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    T s = T(0);
    for (int r = 0; r < world; ++r) s += reinterpret_cast<const volatile T*>(mine)[(size_t)r * nmax + i];
    out[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *epoch_ctr = epoch;
}

}  // namespace

using namespace ptk;

extern "C" {

size_t ptk_allreduce_oneshot_buffer_bytes(int world, int64_t nmax, int itemsize) {
  return (size_t)2 * world * nmax * itemsize + (size_t)2 * world * sizeof(unsigned int) + 64;
}

ptk_status ptk_allreduce_oneshot(int dtype, const void* in, void* out, int64_t n, const uint64_t* peer_ptrs, int rank, int world,
                                 int64_t nmax, void* epoch_ctr, void* stream) {
  PTK_REQUIRE_INIT();
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) return fail(PTK_ERR_ARG, "ptk_allreduce_oneshot: bad rank/world");
  if (n < 0 || n > nmax) return fail(PTK_ERR_ARG, "ptk_allreduce_oneshot: n exceeds the symmetric buffer's nmax");
  if (n == 0) return PTK_OK;
  Peers peers;
  for (int i = 0; i < kMaxWorld; ++i) peers.buf[i] = i < world ? reinterpret_cast<void*>(peer_ptrs[i]) : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PTK_F32)
    allreduce_oneshot_kernel<float><<<1, 256, 0, st>>>((const float*)in, (float*)out, n, peers, rank, world, nmax,
                                                       (unsigned int*)epoch_ctr);
  else if (dtype == PTK_F64)
    allreduce_oneshot_kernel<double><<<1, 256, 0, st>>>((const double*)in, (double*)out, n, peers, rank, world, nmax,
                                                        (unsigned int*)epoch_ctr);
  else
    return fail(PTK_ERR_UNSUPPORTED, "ptk_allreduce_oneshot: dtype must be float32 or float64");
  PTK_LAUNCH_CHECK("allreduce_oneshot");
  return PTK_OK;
}

}  // extern "C"
