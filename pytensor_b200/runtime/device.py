"""Device-buffer plumbing.  torch.cuda tensors are used ONLY as containers (allocation through torch's caching
allocator, shape/stride/dtype bookkeeping, views) — never for arithmetic: every byte that moves or changes on the
device does so in a libptk kernel or a cudaMemcpy issued through the C-ABI.
"""

from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib as _lib

NP_TO_TORCH = {
    "bool": torch.bool, "int8": torch.int8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64,
    "uint8": torch.uint8, "uint16": torch.uint16, "uint32": torch.uint32, "uint64": torch.uint64,
    "float16": torch.float16, "float32": torch.float32, "float64": torch.float64,
}
TORCH_TO_NP = {v: k for k, v in NP_TO_TORCH.items()}


def require_cuda() -> None:
    if _lib.TRACE_ONLY:
        return
    if not torch.cuda.is_available():
        raise _lib.PtkError("the CUDA backend needs a B200 (torch.cuda.is_available() is False); there is no CPU fallback")


_device = None


def device() -> torch.device:
    global _device
    if _lib.TRACE_ONLY:
        return torch.device("meta")
    if _device is None:
        require_cuda()
        _lib.init(torch.cuda.current_device())
        _device = torch.device("cuda", torch.cuda.current_device())
    return _device


def stream_ptr() -> int:
    if _lib.TRACE_ONLY:
        return 0
    return torch.cuda.current_stream().cuda_stream


def is_dev(x) -> bool:
    return isinstance(x, torch.Tensor)


def np_dtype_name(x) -> str:
    if isinstance(x, torch.Tensor):
        return TORCH_TO_NP[x.dtype]
    return np.asarray(x).dtype.name


def empty(shape, dtype: str) -> torch.Tensor:
    return torch.empty(tuple(int(s) for s in shape), dtype=NP_TO_TORCH[dtype], device=device())


def empty_like_layout(shape, dtype: str, order) -> torch.Tensor:
    """Dense buffer whose dims are laid out in `order` (a permutation, slowest first); C order = range(ndim)."""
    shape = tuple(int(s) for s in shape)
    if order is None or tuple(order) == tuple(range(len(shape))):
        return empty(shape, dtype)
    perm_shape = [shape[i] for i in order]
    base = empty(perm_shape, dtype)
    inv = [0] * len(order)
    for pos, dim in enumerate(order):
        inv[dim] = pos
    return base.permute(inv)


def ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def i64_array(values):
    n = max(1, len(values))
    arr = (ctypes.c_int64 * n)(*[int(v) for v in values])
    return arr


def copy_strided(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[...] = src[...] (same shape; src may carry 0 strides for broadcast) with the libptk copy kernel."""
    assert dst.dtype == src.dtype and tuple(dst.shape) == tuple(src.shape), (dst.shape, src.shape)
    if dst.numel() == 0:
        return
    L = _lib.lib()
    nd = dst.dim()
    _lib.check(
        L.ptk_copy_strided(ptr(dst), i64_array(dst.stride()), ptr(src), i64_array(src.stride()),
                           i64_array(dst.shape), nd, dst.element_size(), stream_ptr()),
        "ptk_copy_strided",
    )


def contiguous(t: torch.Tensor) -> torch.Tensor:
    """C-contiguous version of `t` (itself when already contiguous); the copy runs in the libptk copy kernel."""
    if t.is_contiguous():
        return t
    out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    copy_strided(out, t)
    return out


def clone(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    copy_strided(out, t)
    return out


def to_device(arr) -> torch.Tensor:
    """Host -> device copy (cudaMemcpyAsync on the current stream; pinned sources stay asynchronous)."""
    if isinstance(arr, torch.Tensor):
        if arr.is_cuda or arr.is_meta:
            return arr
        arr = arr.numpy()
    a = np.asarray(arr)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    out = empty(a.shape, a.dtype.name)
    if a.size:
        _lib.check(_lib.lib().ptk_memcpy_h2d_async(ptr(out), a.ctypes.data, a.nbytes, stream_ptr()), "h2d")
        # The source may be pageable and freed/mutated by the caller right after we return: make the copy complete.
        # (cudaMemcpyAsync from pageable memory has already staged it; for pinned memory we must wait.)
        _lib.check(_lib.lib().ptk_sync_stream(stream_ptr()), "sync")
    return out


def to_device_async(arr: np.ndarray, out: torch.Tensor | None = None) -> torch.Tensor:
    """H2D without the trailing synchronise: the caller guarantees `arr` outlives the copy (e.g. pinned staging)."""
    a = np.asarray(arr)
    if not a.flags.c_contiguous:
        a = np.ascontiguousarray(a)
    if out is None:
        out = empty(a.shape, a.dtype.name)
    if a.size:
        _lib.check(_lib.lib().ptk_memcpy_h2d_async(ptr(out), a.ctypes.data, a.nbytes, stream_ptr()), "h2d")
    return out


def to_host(t, sync: bool = True) -> np.ndarray:
    """Device -> host copy into a fresh numpy array (C order)."""
    if not isinstance(t, torch.Tensor):
        return np.asarray(t)
    if t.is_meta:
        return np.zeros(tuple(t.shape), dtype=TORCH_TO_NP[t.dtype])
    src = contiguous(t)
    out = np.empty(tuple(src.shape), dtype=TORCH_TO_NP[src.dtype])
    if out.size:
        _lib.check(_lib.lib().ptk_memcpy_d2h_async(out.ctypes.data, ptr(src), out.nbytes, stream_ptr()), "d2h")
    if sync:
        _lib.check(_lib.lib().ptk_sync_stream(stream_ptr()), "sync")
    return out


def synchronize() -> None:
    _lib.check(_lib.lib().ptk_sync_stream(stream_ptr()), "sync")
