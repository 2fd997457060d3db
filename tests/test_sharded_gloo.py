"""CPU suite: the N>1 path (batch sharding + one packed all-reduce) with world_size 2 over gloo.
The per-rank evaluator here is the reference C linker (the oracle) — the host-side sharding logic is what is tested."""

import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {repo!r})
import torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
from oracle import cvm
pytensor = cvm.configure("float64")
from pytensor_b200 import workloads as W
from pytensor_b200.sharded import ShardedSum, shard_bounds
ins, outs, make_args, meta = W.cfg5_logp_grad(B=37, n=24, J=6, K=3, dtype="float64")
f = pytensor.function(ins, outs, mode="CVM")
args = make_args()
sh = ShardedSum(f, batch_arg_idx=[0, 1, 2, 3])
res = sh(*args)
full = f(*args)
# the priors/likelihood are per-chain terms, so shard sums add up exactly to the full-batch sums
for r, e in zip(res, full):
    np.testing.assert_allclose(r, e, rtol=1e-10, atol=1e-10)
lo, hi = shard_bounds(37, 2, dist.get_rank())
assert (hi - lo) in (18, 19)
print("rank", dist.get_rank(), "ok", flush=True)
dist.destroy_process_group()
"""


def test_shard_bounds_cover_exactly():
    sys.path.insert(0, REPO)
    from pytensor_b200.sharded import shard_bounds

    for n in (0, 1, 7, 8, 1 << 20, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_two_rank_gloo_sharded_logp_grad(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=REPO, port=port))
    env = dict(os.environ, PTK_COMPILEDIR=str(tmp_path / "cdir"))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"rank {r} ok" in o
