"""The one-shot NVLink all-reduce kernel (pytensor_b200/csrc/ptk_collective.cu) run on the host with REAL concurrency: every
rank is a forked process (its CTA = OS threads of the kernel emulator), the peer-mapped symmetric buffers are one shared
anonymous mapping visible at the same address in all of them — which is what the NVLink peer pointers are on the device.
Checks the protocol the GPU tests can only exercise at the world sizes a box offers: push / release flag / acquire wait /
ordered sum, the epoch counter in "device" memory, and the two-parity slot reuse under arbitrary skew between ranks."""

import ctypes
import mmap
import multiprocessing as mp
import os
import time

import numpy as np
import pytest

from kernel_emulator import EmulatedKernel

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytensor_b200", "csrc")
MAX_WORLD = 16

SHIM = r"""
#include <sched.h>
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void emu_st_release(unsigned int* p, unsigned int v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline unsigned int emu_ld_acquire(unsigned int* p) { unsigned int v = __atomic_load_n(p, __ATOMIC_ACQUIRE); sched_yield(); return v; }
"""


class Peers(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_void_p * MAX_WORLD)]


def _kernel(tmp_path, ctype):
    text = open(os.path.join(CSRC, "ptk_collective.cu")).read()
    src = text[text.index("constexpr int kMaxWorld"):text.index("}  // namespace")]
    st = 'asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(epoch) : "memory");'
    ld = 'asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");'
    assert st in src and ld in src
    src = src.replace(st, "emu_st_release(f, epoch);").replace(ld, "v = emu_ld_acquire(f);")
    return EmulatedKernel(SHIM + src, "allreduce_oneshot_kernel", tmp_path, threaded=True, template_args=ctype,
                          type_subst={"T": ctype})


def _vec(rank, it, n, dtype):
    rng = np.random.default_rng(1000 * it + rank)
    return (rng.standard_normal(n) * (1 + rank)).astype(dtype)


def _rank_main(k, base, stride, rank, world, nmax, n, iters, dtype, block, q):
    try:
        peers = Peers()
        for r in range(world):
            peers.buf[r] = base + r * stride
        epoch = np.zeros(1, dtype=np.uint32)
        out = np.zeros(n, dtype=dtype)
        rng = np.random.default_rng(rank)
        for it in range(iters):
            if rng.random() < 0.3:
                time.sleep(rng.random() * 0.01)    # skew: this rank arrives late, the others already published
            x = _vec(rank, it, n, dtype)
            k.launch(1, block, [ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(out.ctypes.data), ctypes.c_int64(n), peers,
                                ctypes.c_int(rank), ctypes.c_int(world), ctypes.c_int64(nmax),
                                ctypes.c_void_p(epoch.ctypes.data)])
            want = np.zeros(n, dtype=dtype)
            for r in range(world):   # rank order, like the kernel: bit-identical on every rank
                want = (want + _vec(r, it, n, dtype)).astype(dtype)
            if not np.array_equal(out, want):
                q.put((rank, f"iteration {it}: max diff {np.max(np.abs(out - want))}"))
                return
            if int(epoch[0]) != it + 1:
                q.put((rank, f"iteration {it}: epoch counter {int(epoch[0])}"))
                return
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))


@pytest.mark.parametrize("world,n,dtype,ctype,block", [
    (2, 75, np.float32, "float", 32),      # the logp+grad message of BASELINE configs[4]
    (8, 75, np.float32, "float", 32),      # one NVSwitch domain
    (5, 300, np.float64, "double", 64),    # odd world, n > block, fp64
    (16, 17, np.float32, "float", 32),     # the kernel's maximum world
])
def test_one_shot_all_reduce_protocol_across_processes(tmp_path, world, n, dtype, ctype, block):
    k = _kernel(tmp_path, ctype)
    nmax = 1024
    isz = np.dtype(dtype).itemsize
    stride = (2 * world * nmax * isz + 2 * world * 4 + 64 + 4095) // 4096 * 4096   # ptk_allreduce_oneshot_buffer_bytes, page-rounded
    mm = mmap.mmap(-1, stride * world, flags=mmap.MAP_SHARED | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    iters = 40 if world <= 8 else 15
    procs = [ctx.Process(target=_rank_main, args=(k, base, stride, r, world, nmax, n, iters, dtype, block, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    deadline = time.time() + 240
    try:
        while len(results) < world and time.time() < deadline:
            try:
                r, msg = q.get(timeout=1.0)
                results[r] = msg
            except Exception:  # noqa: BLE001  (queue.Empty)
                if any(p.exitcode not in (None, 0) for p in procs):
                    break
    finally:
        for p in procs:
            p.join(timeout=2)
            if p.is_alive():
                p.kill()   # exactly the processes started above
    assert results == {r: "ok" for r in range(world)}, results
