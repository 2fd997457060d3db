"""cfg2's fused map+row-reduce kernel (K3) on its own: graph-replayed step time over two alternating input sets.
usage: PTK_K3_TPR=64|128|256 python scripts/k3_probe.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import cvm  # noqa: E402

pytensor = cvm.configure("float32")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pytensor_b200  # noqa: E402,F401
from pytensor_b200 import workloads as W  # noqa: E402
from pytensor_b200.link.cuda import cuda_mode  # noqa: E402
from pytensor_b200.runtime import device as dev  # noqa: E402

ins, outs, mk, meta = W.cfg2_fused_elemwise(4096)
f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
sets = [[dev.to_device(a) for a in mk(s)] for s in (1, 101)]
for i in range(8):
    f(*sets[i % 2])
torch.cuda.synchronize()
ts = []
for rep in range(15):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        f(*sets[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50)
ms = float(np.median(ts))
print(f"PTK_K3_TPR={os.environ.get('PTK_K3_TPR', 'auto')}: {ms*1e3:.2f} us/eval  {meta['bytes']/ms/1e6:.0f} GB/s  "
      f"frac of 6479.6 = {meta['bytes']/ms/1e6/6479.6:.3f}  (min {min(ts)*1e3:.2f} us)", flush=True)
