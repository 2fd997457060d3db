#!/usr/bin/env python
"""Round-2 first device probe: (a) plain x+y over (4096,4096) fp32 through mode="CUDA" (streaming efficiency of the K1
vector skeleton on its own), (b) metric graph n=4096 in the default fp32 mode and in bf16 mode."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import cvm  # noqa: E402

pytensor = cvm.configure("float32")
import numpy as np  # noqa: E402
import pytensor.tensor as pt  # noqa: E402
import torch  # noqa: E402

import pytensor_b200  # noqa: E402,F401
from pytensor_b200 import workloads as W  # noqa: E402
from pytensor_b200.link.cuda import cuda_mode  # noqa: E402
from pytensor_b200.runtime import device as dev  # noqa: E402


def time_dev(f, argsets, steps, warm=6):
    for i in range(warm):
        f(*argsets[i % len(argsets)])
    torch.cuda.synchronize()
    ts = []
    for rep in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            f(*argsets[i % len(argsets)])
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / steps)
    return float(np.median(ts)), float(min(ts))


out = {}
x, y = pt.fmatrix("x"), pt.fmatrix("y")
f = pytensor.function([x, y], x + y, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
sets = [[dev.to_device(np.random.default_rng(s + k).standard_normal((4096, 4096)).astype("float32")) for k in range(2)]
        for s in (1, 11)]
med, best = time_dev(f, sets, 50)
out["k1_add_4096"] = {"ms_median": med, "ms_best": best, "GBs": 3 * 4096 * 4096 * 4 / (med * 1e-3) / 1e9}
for n, kw, steps in ((4096, {}, 1), (4096, {"gemm_precision": "bf16"}, 3), (1024, {}, 5), (64, {}, 50)):
    ins, outs, mk, meta = W.metric_graph(n=n)
    f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True, **kw), trust_input=True)
    a = [dev.to_device(v) for v in mk()]
    med, best = time_dev(f, [a], steps, warm=3)
    out[f"metric_n{n}_{kw.get('gemm_precision', 'fp32')}"] = {"ms_median": med, "ms_best": best}
    del f, a
    torch.cuda.empty_cache()
print(json.dumps(out))
