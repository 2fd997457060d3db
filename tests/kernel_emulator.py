"""Test infrastructure: run a GENERATED CUDA kernel on the host, one simulated thread after another.

The elementwise and fused-Scan kernels that `pytensor_b200/codegen` emits contain no barriers, shuffles or shared memory:
every thread is an independent loop nest over `blockIdx/threadIdx/gridDim`.  Compiled as C++ with those built-ins turned
into plain variables, the very same source runs on the CPU, which lets the CPU suite check the kernels' index arithmetic
(grid-stride loops, row/column decomposition, vector tails, circular trace buffers) against NumPy without a GPU.
Vector loads/stores keep their alignment requirement: the shim asserts it, because a misaligned 16-byte access that the
CPU would tolerate is a fault on the device.  (Kernels with warp shuffles / shared memory — the reductions — and the
hand-written libptk kernels are covered by the -m gpu suite only.)"""

import ctypes
import re
import subprocess

from pytensor_b200.codegen.scalar import PRELUDE

HOST_PRELUDE = r"""
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
using std::min; using std::max; using std::isnan; using std::isinf;
#define __device__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
struct emu_dim3 { unsigned x, y, z; };
static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline double __longlong_as_double(long long v) { double f; std::memcpy(&f, &v, 8); return f; }
static inline float ptk_max_nan_f32(float a, float b) { return (b > a) ? b : ((a >= b) ? a : NAN); }
static inline float ptk_min_nan_f32(float a, float b) { return (b < a) ? b : ((a <= b) ? a : NAN); }
template <typename T> static inline T ptk_floordiv(T x, T y) { if (y == 0) return 0; T q = x / y; if ((x % y != 0) && ((x < 0) != (y < 0))) --q; return q; }
template <typename T> static inline T ptk_imod_py(T x, T y) { if (y == 0) return 0; T r = x % y; if (r != 0 && ((r < 0) != (y < 0))) r += y; return r; }
template <typename T> static inline T ptk_fmod_py(T x, T y) { T r = std::fmod(x, y); if (r != 0 && ((r < 0) != (y < 0))) r += y; return r; }
"""

# the vector helpers of codegen/elemwise.py, with the device's alignment rule made explicit
ALIGN_CHECKED_VEC = r"""
template <typename T, int N> struct alignas(sizeof(T) * N) PVec { T v[N]; };
static inline void emu_check_align(const void* p, size_t a) {
  if (reinterpret_cast<uintptr_t>(p) % a) { std::fprintf(stderr, "misaligned %zu-byte vector access\n", a); std::abort(); }
}
template <typename T, int N> static inline PVec<T, N> ptk_ldv(const T* p) {
  emu_check_align(p, sizeof(T) * N); PVec<T, N> r; std::memcpy(&r, p, sizeof(r)); return r;
}
template <typename T, int N> static inline void ptk_stv(T* p, const PVec<T, N>& v) {
  emu_check_align(p, sizeof(T) * N); std::memcpy(p, &v, sizeof(v));
}
template <typename T, int N> static inline PVec<T, N> ptk_ldv_pin(const T* p) { return ptk_ldv<T, N>(p); }
static inline void ptk_prefetch_l2(const void*) {}
"""


# Host stand-ins for the mbarrier / bulk-copy wrappers of the TMA-staged row kernel (codegen/careduce.py _TMA_HELPERS): one
# 64-bit word per barrier — bit 63 phase, bits 40..55 arrival count, bits 24..39 pending arrivals, bits 0..23 pending bytes.
EMU_TMA = r"""
static inline bool emu_bar_update(unsigned long long* bar, long long d_pending, long long d_tx) {
  unsigned long long old = __atomic_load_n(bar, __ATOMIC_SEQ_CST), neu;
  do {
    unsigned long long phase = old >> 63, count = (old >> 40) & 0xffffull;
    long long pending = (long long)((old >> 24) & 0xffffull) + d_pending, tx = (long long)(old & 0xffffffull) + d_tx;
    if (pending < 0 || tx < 0 || tx > 0xffffff) { std::fprintf(stderr, "mbarrier protocol violated\n"); std::abort(); }
    if (pending == 0 && tx == 0) { phase ^= 1ull; pending = (long long)count; }
    neu = (phase << 63) | (count << 40) | ((unsigned long long)pending << 24) | (unsigned long long)tx;
  } while (!__atomic_compare_exchange_n(bar, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return true;
}
static inline void ptk_mbar_init(unsigned long long* bar, unsigned count) {
  __atomic_store_n(bar, ((unsigned long long)count << 40) | ((unsigned long long)count << 24), __ATOMIC_SEQ_CST);
}
static inline void ptk_mbar_fence_init() {}
static inline void ptk_mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu_bar_update(bar, -1, (long long)bytes); }
static inline void ptk_mbar_arrive(unsigned long long* bar) { emu_bar_update(bar, -1, 0); }
static inline void ptk_mbar_wait(unsigned long long* bar, unsigned parity) {
  while ((unsigned)(__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 63) == parity) std::this_thread::yield();
}
static inline void ptk_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src) | bytes) & 15u) {
    std::fprintf(stderr, "bulk copy needs 16-byte aligned addresses and size\n"); std::abort();
  }
  std::memcpy(dst, src, bytes);
  emu_bar_update(bar, 0, -(long long)bytes);
}
"""


# Threaded variant for kernels with warp shuffles, __syncthreads and __shared__ memory (the row reductions): every
# simulated thread of a block is a real OS thread, a shuffle is "publish my value, warp barrier, read the partner's value,
# warp barrier", __syncthreads is a block barrier, `__shared__` becomes a function-local static (blocks run one at a time).
HOST_PRELUDE_MT = HOST_PRELUDE.replace("static emu_dim3 threadIdx, blockIdx, blockDim, gridDim;",
                                       "static thread_local emu_dim3 threadIdx;\nstatic emu_dim3 blockIdx, blockDim, gridDim;") + r"""
#include <pthread.h>
#include <thread>
#include <vector>
#define __shared__ static
static pthread_barrier_t emu_block_bar;
static pthread_barrier_t emu_warp_bar[32];
static unsigned long long emu_slot[1024];
static inline void __syncthreads() { pthread_barrier_wait(&emu_block_bar); }
template <typename T> static inline T emu_exchange(T v, int src_lane_in_block) {
  const int tid = threadIdx.x, w = tid >> 5;
  std::memcpy(&emu_slot[tid], &v, sizeof(T));
  pthread_barrier_wait(&emu_warp_bar[w]);
  T r; std::memcpy(&r, &emu_slot[src_lane_in_block], sizeof(T));
  pthread_barrier_wait(&emu_warp_bar[w]);
  return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)(threadIdx.x ^ m)); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31; return emu_exchange(v, lane + d < 32 ? (int)threadIdx.x + d : (int)threadIdx.x);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x & 31; return emu_exchange(v, lane - d >= 0 ? (int)threadIdx.x - d : (int)threadIdx.x);
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
"""


def extract_static_kernel(cu_path: str, name: str) -> str:
    """Source text of one `__global__` (template) kernel of a hand-written .cu file: from its `template <...>` line (if any)
    to the closing brace in column 0."""
    text = open(cu_path).read()
    m = re.search(r"^(?:template <[^\n]*>\n)?__global__ void[^\n]*\b" + re.escape(name) + r"\(", text, re.M)
    assert m, f"{name} not found in {cu_path}"
    end = text.index("\n}\n", m.start()) + 3
    return text[m.start():end]


def _strip_launch_bounds(text: str) -> str:
    out, pos = [], 0
    while True:
        i = text.find("__launch_bounds__(", pos)
        if i < 0:
            out.append(text[pos:])
            return "".join(out)
        out.append(text[pos:i])
        depth, j = 0, i + len("__launch_bounds__")
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            j += 1
            if depth == 0:
                break
        pos = j


STATIC_SHIM = r"""
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline void __syncwarp() {}
"""

# Warp-cooperative generated kernels (codegen/rowfuse.py): lanes hand values over through shared memory, so __syncwarp is
# a REAL warp barrier here; shared/global atomics are host atomics; __ldg is a plain load; 128-bit vector types.
WARP_SHIM = r"""
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline void __syncwarp() { pthread_barrier_wait(&emu_warp_bar[threadIdx.x >> 5]); }
static inline float atomicAdd(float* p, float v) {
  unsigned int* u = reinterpret_cast<unsigned int*>(p); unsigned int old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
  float f;
  do { std::memcpy(&f, &old, 4); f += v; std::memcpy(&neu, &f, 4); }
  while (!__atomic_compare_exchange_n(u, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return f - v;
}
static inline double atomicAdd(double* p, double v) {
  unsigned long long* u = reinterpret_cast<unsigned long long*>(p); unsigned long long old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
  double f;
  do { std::memcpy(&f, &old, 8); f += v; std::memcpy(&neu, &f, 8); }
  while (!__atomic_compare_exchange_n(u, &old, neu, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return f - v;
}
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
using std::fma;
"""


class EmulatedKernel:
    def __init__(self, source: str, name: str, tmp_path, threaded: bool = False, template_args: str = "",
                 type_subst: dict | None = None, dynamic_smem: str | None = None, warp_shim: bool = False):
        """`template_args` / `type_subst`: instantiate a template kernel (e.g. "float, 8", {"T": "float"});
        `dynamic_smem`: name of the kernel's `extern __shared__` array (gets a fixed 100 KiB static buffer)."""
        from pytensor_b200.codegen.elemwise import _VEC_HELPERS

        from pytensor_b200.codegen.careduce import _TMA_HELPERS

        body = source.replace(PRELUDE, "").replace(_VEC_HELPERS, ALIGN_CHECKED_VEC).replace(_TMA_HELPERS, EMU_TMA)
        body = body.replace("__shared__ __align__(128)", "alignas(128) static")
        if dynamic_smem:
            body = re.sub(r"extern __shared__[^\n;]*\b" + re.escape(dynamic_smem) + r"\[\];",
                          f"alignas(16) static unsigned char {dynamic_smem}[160 * 1024];", body)
            body = (WARP_SHIM if warp_shim else STATIC_SHIM) + body
        elif warp_shim:
            body = WARP_SHIM + body
        body = _strip_launch_bounds(body)
        m = re.search(r"__global__ void\s+" + re.escape(name) + r"\((.*?)\) \{", body, re.S)
        assert m, "kernel signature not found"
        self.param_types = []
        for p in m.group(1).split(","):
            p = " ".join(p.split())
            t = p[: re.search(r"[A-Za-z_0-9]+$", p).start()].strip()
            for a, b in (type_subst or {}).items():
                t = re.sub(r"\b" + a + r"\b", b, t)
            self.param_types.append(t)
        if template_args:
            name_call = f"{name}<{template_args}>"
        else:
            name_call = name
        unpack = ", ".join(f"*reinterpret_cast<{t.replace('const ', '', 1) if not t.endswith('*') else t}*>(a[{k}])"
                           for k, t in enumerate(self.param_types))
        if not threaded:
            wrapper = (f'\nextern "C" void emu_launch(unsigned gx, unsigned gy, unsigned gz, unsigned block, void** a) {{\n'
                       f"  gridDim = {{gx, gy, gz}}; blockDim = {{block, 1, 1}};\n"
                       f"  for (unsigned bz = 0; bz < gz; ++bz) for (unsigned by = 0; by < gy; ++by) for (unsigned b = 0; b < gx; ++b)\n"
                       f"    for (unsigned t = 0; t < block; ++t) {{\n"
                       f"      blockIdx = {{b, by, bz}}; threadIdx = {{t, 0, 0}};\n      {name_call}({unpack});\n    }}\n}}\n")
        else:
            wrapper = (f'\nextern "C" void emu_launch(unsigned gx, unsigned gy, unsigned gz, unsigned block, void** a) {{\n'
                       f"  gridDim = {{gx, gy, gz}}; blockDim = {{block, 1, 1}};\n"
                       f"  pthread_barrier_init(&emu_block_bar, nullptr, block);\n"
                       f"  for (unsigned w = 0; w < block / 32; ++w) pthread_barrier_init(&emu_warp_bar[w], nullptr, 32);\n"
                       f"  for (unsigned bz = 0; bz < gz; ++bz) for (unsigned by = 0; by < gy; ++by) for (unsigned b = 0; b < gx; ++b) {{\n"
                       f"    blockIdx = {{b, by, bz}};\n    std::vector<std::thread> ts;\n"
                       f"    for (unsigned t = 0; t < block; ++t) ts.emplace_back([=]() {{ threadIdx = {{t, 0, 0}}; {name_call}({unpack}); }});\n"
                       f"    for (auto& th : ts) th.join();\n  }}\n}}\n")
        tag = re.sub(r"[^A-Za-z0-9]+", "_", name_call)
        cpp, so = tmp_path / f"{tag}.cpp", tmp_path / f"{tag}.so"
        cpp.write_text((HOST_PRELUDE_MT if threaded else HOST_PRELUDE) + body + wrapper)
        subprocess.run(["g++", "-O1", "-march=native", "-fno-math-errno", "-shared", "-fPIC", "-std=c++17", "-w", "-pthread",
                        str(cpp), "-o", str(so)], check=True)
        self.lib = ctypes.CDLL(str(so))

    def launch(self, grid, block: int, args):
        """`grid`: int or (x, y, z); `args`: ctypes objects in kernel-parameter order (pointers = c_void_p to HOST memory)."""
        assert len(args) == len(self.param_types), (len(args), self.param_types)
        g = (grid, 1, 1) if isinstance(grid, int) else tuple(grid) + (1,) * (3 - len(grid))
        assert block % 32 == 0 or not hasattr(self.lib, "dummy")
        arr = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
        self.lib.emu_launch(ctypes.c_uint(g[0]), ctypes.c_uint(g[1]), ctypes.c_uint(g[2]), ctypes.c_uint(block), arr)
