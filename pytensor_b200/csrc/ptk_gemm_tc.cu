// placeholder: replaced by the tcgen05/TMEM kernel
#include "ptk_common.h"
namespace ptk {
size_t gemm_tc_workspace(int64_t M, int64_t N, int64_t K) { return 0; }
ptk_status gemm_tc(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sa0, int64_t sa1,
                   const float* B, int64_t sb0, int64_t sb1, float beta, float* C, int64_t sc0, int64_t sc1,
                   const float* bias, int act, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  return fail(PTK_ERR_UNSUPPORTED, "gemm_tc not built");
}
}
