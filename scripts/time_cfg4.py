"""Time BASELINE configs[3] (Scan, 1000 steps, state (8192,512) fp32): final-state-only and full-trace variants."""
import sys

import torch

sys.path.insert(0, ".")
import bench
from oracle import cvm

pytensor = cvm.configure("float32")
import pytensor_b200  # noqa: F401
from pytensor_b200 import workloads as W
from pytensor_b200.link.cuda import cuda_mode
from pytensor_b200.runtime import device as dev

for full in (False, True):
    ins, outs, make_args, meta = W.cfg4_scan(8192, 512, 1000, full_trace=full)
    f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
    a = [dev.to_device(x) for x in make_args()]
    ms = bench._time_dev(f, a, torch, 4, 3)
    print("cfg4", "full trace" if full else "final state", "ms", round(ms, 3), "trace GB/s", round(meta["bytes"] / ms / 1e6, 1), flush=True)
    del f, a
    torch.cuda.empty_cache()
