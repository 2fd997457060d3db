"""Multi-GPU: the one-shot NVLink all-reduce kernel and the sharded logp+grad evaluation against NCCL and a
single-process evaluation of the full batch.  Needs >= 2 GPUs (skipped on a 1-GPU box)."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {repo!r})
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from oracle import cvm
pytensor = cvm.configure("float32")
import pytensor_b200
from pytensor_b200 import workloads as W
from pytensor_b200.link.cuda import cuda_mode
from pytensor_b200.runtime import device as dev
from pytensor_b200.sharded import PeerAllReduce, ShardedSum
dev.device()
ar = PeerAllReduce(nmax=1024, dtype="float32")
for it in range(7):                       # many epochs: exercises both parities of the symmetric buffer
    n = 75 if it % 2 == 0 else 1024
    x = dev.to_device((np.arange(n, dtype="float32") + 1) * (rank + 1) * (it + 1))
    ref = x.clone(); dist.all_reduce(ref)
    got = dev.to_host(ar(x))
    np.testing.assert_allclose(got, dev.to_host(ref), rtol=1e-6)
    np.testing.assert_allclose(got, (np.arange(n) + 1) * (it + 1) * sum(r + 1 for r in range(world)), rtol=1e-6)
# sharded logp+grad: one-shot and NCCL agree with each other and with the full batch on one rank
B = 512 * world
ins, outs, mk, meta = W.cfg5_logp_grad(B=B, n=128, J=16, K=4, packed=True)
f = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True, borrow_outputs=True), trust_input=True)
full = mk(seed=5)
lo, hi = rank * 512, (rank + 1) * 512
local = [dev.to_device(a[lo:hi]) if i < 4 else dev.to_device(a) for i, a in enumerate(full)]
res = {{}}
for coll in ("nccl", "oneshot"):
    sh = ShardedSum(f, [0, 1, 2, 3], collective=coll)
    for _ in range(4):
        out = sh(*local, presharded=True)
    res[coll] = dev.to_host(out[0]).copy()
np.testing.assert_allclose(res["nccl"], res["oneshot"], rtol=2e-5, atol=1e-3)
f_full = pytensor.function(ins, outs, mode="CUDA")
whole = f_full(*full)[0]
np.testing.assert_allclose(res["oneshot"], whole, rtol=2e-4, atol=2e-2)
print("rank", rank, "ok", flush=True)
dist.destroy_process_group()
"""


def test_oneshot_allreduce_and_sharded_logp(gpu, tmp_path):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO))
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, PYTHONPATH=REPO))
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    for r in range(world):
        assert f"rank {r} ok" in p.stdout
