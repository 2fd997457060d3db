"""Fixed cost of one tensor-core GEMM launch: staged operands, small problems, back-to-back launches (CUDA events), eager
and as nodes of one CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytensor_b200.runtime import device as dev
from pytensor_b200.runtime import lib as _lib
from pytensor_b200.vm import nodes_blas as nb

L = _lib.init(0)
dev.device()
for (M, N, K) in [(256, 256, 256), (1024, 1024, 1024), (8192, 512, 512), (4096, 4096, 512), (4096, 4096, 4096)]:
    for pieces, terms in ((1, 1), (3, 6)):
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(K, N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        Ast = nb.stage_operand(A, pieces)
        Bst = nb.stage_operand(B, pieces, transposed=True)

        def run():
            nb.gemm_staged(Ast, Bst, terms, 1.0, 0.0, C)

        for _ in range(5):
            run()
        torch.cuda.synchronize()
        n = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / n
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                run()
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / n
        print(f"M={M} N={N} K={K} terms={terms}: eager {eager*1e3:7.1f} us/launch   graph {graph*1e3:7.1f} us/node   "
              f"({2*M*N*K*max(1,terms)/graph/1e9:6.0f} TF/s bf16-equiv in graph)", flush=True)
