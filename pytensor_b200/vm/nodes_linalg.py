"""Cholesky / SolveTriangular nodes (reference: pytensor/tensor/linalg/decomposition/cholesky.py:18,
solvers/triangular.py:13; batched through Blockwise, tensor/blockwise.py:153)."""

from __future__ import annotations

from ..runtime import device as dev
from ..runtime import lib as _lib
from .nodes_elemwise import Node
from .values import Val


class CholeskyNode(Node):
    serial_group = "linalg"  # the blocked kernels share one device status word: never overlap two of them

    def __init__(self, dtype, lower=True, name="Cholesky"):
        self.dtype, self.lower, self.name = dtype, lower, name

    def run(self, vals):
        a = vals[0].dev()
        if a.shape[-1] != a.shape[-2]:
            raise ValueError("Cholesky: last two dims must be square")
        out = dev.clone(a) if a.numel() else dev.empty(tuple(a.shape), self.dtype)
        n = a.shape[-1]
        batch = out.numel() // (n * n) if n else 0
        if n and batch:
            _lib.check(_lib.lib().ptk_potrf(_lib.DTYPE_CODE[self.dtype], dev.ptr(out), n, batch, 1 if self.lower else 0,
                                            dev.stream_ptr()), "ptk_potrf")
        return [Val(d=out)]


class SolveTriangularNode(Node):
    """x = solve(op(A), b): lower/upper, trans, unit_diagonal, b_ndim 1|2 (triangular.py:16-21)."""

    serial_group = "linalg"

    def __init__(self, dtype, lower, unit_diagonal, b_ndim, trans=0, name="SolveTriangular"):
        self.dtype, self.lower, self.unit_diagonal, self.b_ndim, self.trans, self.name = (
            dtype, lower, unit_diagonal, b_ndim, trans, name)

    def run(self, vals):
        A = dev.contiguous(vals[0].dev())
        b = vals[1].dev()
        n = A.shape[-1]
        if self.b_ndim == 1:
            bb = b.unsqueeze(-1)
        else:
            bb = b
        if bb.shape[-2] != n:
            raise ValueError("SolveTriangular: A and b have incompatible shapes")
        out = dev.clone(bb) if bb.numel() else dev.empty(tuple(bb.shape), self.dtype)
        nrhs = bb.shape[-1]
        batchA = A.numel() // (n * n) if n else 0
        batchB = out.numel() // (n * nrhs) if (n and nrhs) else 0
        if n and nrhs and batchB:
            if batchA != batchB:
                raise NotImplementedError("SolveTriangular: broadcasting between batched A and b")
            _lib.check(_lib.lib().ptk_trsm(_lib.DTYPE_CODE[self.dtype], dev.ptr(A), dev.ptr(out), n, nrhs, batchB,
                                           1 if self.lower else 0, 1 if self.trans else 0,
                                           1 if self.unit_diagonal else 0, dev.stream_ptr()), "ptk_trsm")
        if self.b_ndim == 1:
            out = out.squeeze(-1)
        return [Val(d=out)]
