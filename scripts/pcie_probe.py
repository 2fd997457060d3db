"""PCIe probe: raw pinned H2D / D2H bandwidth through the C-ABI copies, alone and concurrently on two streams, and a
cProfile of one host-buffer call of cfg2 (where does the e2e time go?)."""
import cProfile
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import cvm

pytensor = cvm.configure("float32")
import bench
import pytensor_b200
from pytensor_b200 import workloads as W
from pytensor_b200.runtime import device as dev, lib as _lib

L = _lib.init(0)
n = 64 << 20
h_in = bench.pinned_like(np.zeros(n, np.uint8))
h_out = bench.pinned_like(np.zeros(n, np.uint8))
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps


def h2d():
    L.ptk_memcpy_h2d_async(d_a.data_ptr(), h_in.ctypes.data, n, s1.cuda_stream)


def d2h():
    L.ptk_memcpy_d2h_async(h_out.ctypes.data, d_b.data_ptr(), n, s2.cuda_stream)


def both():
    h2d()
    d2h()


t = timed(h2d); print(f"H2D 64 MiB pinned: {t*1e3:.3f} ms  {n/t/1e9:.1f} GB/s")
t = timed(d2h); print(f"D2H 64 MiB pinned: {t*1e3:.3f} ms  {n/t/1e9:.1f} GB/s")
t = timed(both); print(f"H2D + D2H concurrently (2 streams): {t*1e3:.3f} ms  {2*n/t/1e9:.1f} GB/s aggregate")
pg = np.zeros(n, np.uint8)
def h2d_pageable():
    L.ptk_memcpy_h2d_async(d_a.data_ptr(), pg.ctypes.data, n, s1.cuda_stream)
t = timed(h2d_pageable); print(f"H2D 64 MiB pageable: {t*1e3:.3f} ms  {n/t/1e9:.1f} GB/s")

ins, outs, mk, _ = W.cfg2_fused_elemwise(4096)
f = pytensor.function(ins, outs, mode="CUDA", trust_input=True)
args = [bench.pinned_like(a) for a in mk(1)]
for _ in range(3):
    f(*args)
ex = f.vm.executor
for chunks, minb in ((0, 1 << 60), (4, 0), (8, 0), (16, 0), (32, 0)):
    ex.STREAM_MIN_BYTES = minb
    ex.STREAM_CHUNKS = max(chunks, 2)
    ex.STREAM_CHUNK_BYTES = 1 << 20
    ex._chunk_plan = None
    for _ in range(4):
        r = f(*args)
    ts = []
    for _ in range(15):
        t = time.perf_counter()
        r = f(*args)
        ts.append(time.perf_counter() - t)
    print("one-shot" if chunks == 0 else f"chunks={chunks}", "e2e ms median", np.median(ts) * 1e3, "min", min(ts) * 1e3,
          "max", max(ts) * 1e3, flush=True)
pg_args = mk(1)
ex.STREAM_CHUNKS = 16
ex._chunk_plan = None
for label, minb in (("pageable one-shot", 1 << 60), ("pageable chunks=16", 0)):
    ex.STREAM_MIN_BYTES = minb
    for _ in range(3):
        r = f(*pg_args)
    t = time.perf_counter()
    for _ in range(10):
        r = f(*pg_args)
    print(label, "e2e ms", (time.perf_counter() - t) / 10 * 1e3, flush=True)
