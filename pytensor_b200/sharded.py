"""Batch-sharded evaluation across GPUs (BASELINE.json configs[4]).

Independent evaluations (PyMC-style logp+grad over chains) are split along the batch axis into `world` contiguous
shards, one process per GPU; every rank runs the SAME compiled function on its shard and produces batch-summed partials
`[logp, grads...]`; ONE all-reduce (sum) of the packed partials makes the result valid on every rank.  The reference has
no distributed code at all, so there is no reference interface to mirror: this is host-side plumbing
over `torch.distributed` (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra item."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_args(args, batch_arg_idx, world, rank):
    """Slice the batch-axis (dim 0) arguments for this rank; replicate the others."""
    out = list(args)
    n = None
    for i in batch_arg_idx:
        ni = args[i].shape[0]
        if n is None:
            n = ni
        elif ni != n:
            raise ValueError("batch arguments disagree on the batch size")
        lo, hi = shard_bounds(ni, world, rank)
        out[i] = args[i][lo:hi]
    return out


class PeerAllReduce:
    """One-shot all-reduce of a small vector over NVLink peer memory (libptk `ptk_allreduce_oneshot`).

    torch.distributed's symmetric-memory rendezvous is used ONLY to obtain peer-mapped pointers to every rank's buffer;
    the data movement and the reduction are one hand-written kernel per rank (push to all peers, flag, wait, sum)."""

    def __init__(self, group=None, nmax=1024, dtype="float32"):
        import ctypes

        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        from pytensor_b200.runtime import device as dev
        from pytensor_b200.runtime import lib as _lib

        self.dist = dist
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.nmax = int(nmax)
        self.dtype = dtype
        isz = 4 if dtype == "float32" else 8
        L = _lib.lib()
        nbytes = int(L.ptk_allreduce_oneshot_buffer_bytes(self.world, self.nmax, isz))
        dev.device()
        self.buf = symm.empty((nbytes + 3) // 4, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        _lib.check(L.ptk_memset_async(self.buf.data_ptr(), 0, self.buf.numel() * 4, dev.stream_ptr()), "memset")
        dev.synchronize()
        try:  # needed by older torch releases, a no-op / deprecated in newer ones
            import warnings

            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001
            pass
        self.handle = symm.rendezvous(self.buf, self.group.group_name)
        ptrs = list(self.handle.buffer_ptrs)
        self.peer_ptrs = (ctypes.c_uint64 * self.world)(*[int(p) for p in ptrs])
        self.epoch = dev.empty((1,), "int32")
        _lib.check(L.ptk_memset_async(self.epoch.data_ptr(), 0, 4, dev.stream_ptr()), "memset")
        dev.synchronize()
        dist.barrier(self.group)  # every rank's buffer is zeroed before the first push can land
        self._L, self._lib, self._dev = L, _lib, dev

    def __call__(self, x, out=None):
        """x: contiguous device vector (n <= nmax). Returns the element-wise sum over ranks (in `out` or in place)."""
        n = x.numel()
        if out is None:
            out = x
        self._lib.check(
            self._L.ptk_allreduce_oneshot(self._lib.DTYPE_CODE[self.dtype], x.data_ptr(), out.data_ptr(), n, self.peer_ptrs,
                                          self.rank, self.world, self.nmax, self.epoch.data_ptr(), self._dev.stream_ptr()),
            "ptk_allreduce_oneshot")
        return out


class ShardedSum:
    """Callable wrapper: `f_local(*local_args) -> list of batch-summed partial outputs` on each rank, then one packed
    all-reduce.  Works for NumPy outputs (CPU/gloo) and device tensors (NCCL)."""

    def __init__(self, f_local, batch_arg_idx, group=None, collective="nccl"):
        """collective: "nccl" (torch.distributed all_reduce after the call), "oneshot" (PeerAllReduce kernel launched
        after the call; packed device outputs only), "oneshot_ingraph" (the same kernel registered as the executor's
        epilogue: it is part of the function's launch list, hence of its captured CUDA graph — one cudaGraphLaunch per
        evaluation covers compute AND the all-reduce; packed device output only), "none" (single process)."""
        import torch.distributed as dist

        self.collective = collective
        self._peer = None
        self._ingraph = False
        self.f = f_local
        self.batch_arg_idx = list(batch_arg_idx)
        self.group = group
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._buf = None

    def local_args(self, args):
        return shard_args(args, self.batch_arg_idx, self.world, self.rank)

    def describe(self):
        if self.world == 1:
            return None
        return {"nccl": "ncclAllReduce (torch.distributed) after the graph replay",
                "oneshot": "hand-written one-shot NVLink all-reduce kernel after the graph replay",
                "oneshot_ingraph": "hand-written one-shot NVLink all-reduce kernel INSIDE the captured CUDA graph"
                }.get(self.collective, self.collective) + ("" if self.collective != "oneshot_ingraph" or self._ingraph
                                                           else " (NOT installed: fell back to launching it after the call)")

    def _install_ingraph(self):
        """Register the one-shot all-reduce as the executor's epilogue (runs right after the last node, on the VM
        stream, also while capturing): needs a packed single device output."""
        import torch

        vm = getattr(self.f, "vm", None)
        ex = getattr(vm, "executor", None)
        if ex is None or not getattr(vm, "device_outputs", False):
            return False
        peer_box = {}

        def epilogue(out_vals):
            v = out_vals[0]
            if len(out_vals) != 1 or v.d is None or not v.d.is_contiguous() or v.d.numel() > 1024 \
                    or v.d.dtype not in (torch.float32, torch.float64):
                raise RuntimeError("oneshot_ingraph needs ONE packed contiguous float device output of <= 1024 elements")
            if "p" not in peer_box:
                from pytensor_b200.runtime import device as dev

                if dev.alloc_state.capturing:
                    raise dev.GraphUnsupported("peer rendezvous inside a capture")
                peer_box["p"] = PeerAllReduce(self.group, 1024, "float32" if v.d.dtype == torch.float32 else "float64")
            peer_box["p"](v.d)

        ex.epilogue = epilogue
        self._peer_box = peer_box
        return True

    def reduce(self, outs):
        """Pack -> all_reduce(sum) -> unpack. One collective per evaluation regardless of the number of outputs."""
        import torch

        if self.world == 1:
            return outs
        sizes = [int(np.prod(o.shape)) if len(o.shape) else 1 for o in outs]
        total = sum(sizes)
        if len(outs) == 1 and isinstance(outs[0], torch.Tensor) and outs[0].is_contiguous():
            # the graph already packed its partials into one vector: reduce it in place, ONE collective, no copies
            if self.collective == "oneshot" and outs[0].numel() <= 1024 and outs[0].dtype in (torch.float32, torch.float64):
                if self._peer is None:
                    self._peer = PeerAllReduce(self.group, 1024, "float32" if outs[0].dtype == torch.float32 else "float64")
                self._peer(outs[0])
                return outs
            self.dist.all_reduce(outs[0], op=self.dist.ReduceOp.SUM, group=self.group)
            return outs
        if isinstance(outs[0], torch.Tensor):
            from pytensor_b200.runtime import device as dev

            if self._buf is None or self._buf.numel() != total or self._buf.dtype != outs[0].dtype:
                self._buf = torch.empty((total,), dtype=outs[0].dtype, device=outs[0].device)
            pos = 0
            for o, n in zip(outs, sizes):
                dev.copy_strided(self._buf[pos:pos + n].view(o.shape), o)
                pos += n
            self.dist.all_reduce(self._buf, op=self.dist.ReduceOp.SUM, group=self.group)
            res, pos = [], 0
            for o, n in zip(outs, sizes):
                res.append(self._buf[pos:pos + n].view(o.shape))
                pos += n
            return res
        flat = torch.from_numpy(np.concatenate([np.asarray(o, dtype=np.float64).reshape(-1) for o in outs]))
        self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group)
        res, pos = [], 0
        flat = flat.numpy()
        for o, n in zip(outs, sizes):
            res.append(flat[pos:pos + n].reshape(np.shape(o)).astype(np.asarray(o).dtype))
            pos += n
        return res

    def __call__(self, *args, presharded=False):
        local = list(args) if presharded else self.local_args(args)
        if self.collective == "oneshot_ingraph" and self.world > 1 and not self._ingraph:
            self._ingraph = self._install_ingraph()
            if not self._ingraph:
                self.collective = "oneshot"
        outs = self.f(*local)
        if not isinstance(outs, list | tuple):
            outs = [outs]
        if self._ingraph or self.collective == "none":
            return list(outs)
        return self.reduce(list(outs))
