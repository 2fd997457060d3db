// Internal helpers shared by the libptk translation units (not part of the C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "ptk.h"

namespace ptk {

void set_error(const std::string& msg);
ptk_status fail(ptk_status code, const std::string& msg);
ptk_status check_cuda(cudaError_t e, const char* what);
ptk_status check_cu(CUresult r, const char* what);
int sm_count();

struct DriverApi {
  CUresult (*ModuleLoadDataEx)(CUmodule*, const void*, unsigned, CUjit_option*, void**) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
};
const DriverApi& drv();
bool initialised();

inline int dtype_size(int dt) {
  switch (dt) {
    case PTK_BOOL: case PTK_I8: case PTK_U8: return 1;
    case PTK_I16: case PTK_U16: case PTK_F16: return 2;
    case PTK_I32: case PTK_U32: case PTK_F32: return 4;
    case PTK_I64: case PTK_U64: case PTK_F64: return 8;
  }
  return 0;
}

}  // namespace ptk

#define PTK_REQUIRE_INIT()                                                                      \
  do {                                                                                          \
    if (!ptk::initialised()) return ptk::fail(PTK_ERR_CUDA, "ptk_init() has not been called");  \
  } while (0)
#define PTK_CUDA(expr)                                               \
  do {                                                               \
    ptk_status _s = ptk::check_cuda((expr), #expr);                  \
    if (_s != PTK_OK) return _s;                                     \
  } while (0)
#define PTK_CU(expr)                                                 \
  do {                                                               \
    ptk_status _s = ptk::check_cu((expr), #expr);                    \
    if (_s != PTK_OK) return _s;                                     \
  } while (0)
#define PTK_LAUNCH_CHECK(what) PTK_CUDA(cudaPeekAtLastError())
