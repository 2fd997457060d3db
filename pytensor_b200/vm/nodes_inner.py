"""Nodes that own an inner lowered Program (other than Scan): OpFromGraph (pytensor/compile/builders.py:116)."""

from __future__ import annotations

from .nodes_elemwise import Node


class InnerProgramNode(Node):
    """Runs a lowered inner graph inline: outputs = inner_program(inputs).  The reference calls a nested compiled
    `Function` here (builders.py:878-882); on the device the inner steps are simply part of the same launch stream (and of
    the same captured CUDA graph)."""

    def __init__(self, program, n_out, name="OpFromGraph"):
        from .vm import Executor

        self.inner = Executor(program, allow_gc=True, use_graph=False)
        self.n_out = n_out
        self.name = name

    def run(self, vals):
        return self.inner.run(list(vals))
