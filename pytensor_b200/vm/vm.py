"""The CUDA VM: executes a lowered program (a topologically ordered list of executable nodes) over device buffers.

Role of the reference's `Loop`/`Stack`/CVM (pytensor/link/vm.py:330,415; link/c/c_code/lazylinker_c.c:749-897):
run every thunk in order, drop intermediates after their last use, hand the outputs back, record the failing
position for `raise_with_op`.  Differences that matter on a GPU:
  * values are `Val`s (device buffers; host copies only for inputs/constants/shape integers);
  * input cells are read, never overwritten (so `SharedVariable.get_value` keeps seeing NumPy);
  * outputs are copied device->host once at the end (one synchronisation per call) unless `device_outputs`.
No pytensor import here: a program can be pickled and replayed on a box with only torch + libptk.
"""

from __future__ import annotations

import time

import numpy as np

from ..runtime import device as dev
from . import nodes_basic
from .values import Val, wrap


class Step:
    __slots__ = ("impl", "ins", "outs", "free", "origin")

    def __init__(self, impl, ins, outs, origin=-1):
        self.impl, self.ins, self.outs, self.free, self.origin = impl, list(ins), list(outs), [], origin


class Program:
    """Flat, pytensor-free description of a compiled graph.

    n_slots value slots; `inputs` / `outputs` are slot ids; `constants` maps slot -> numpy array."""

    def __init__(self, n_slots, inputs, outputs, constants, steps):
        self.n_slots = n_slots
        self.inputs = list(inputs)
        self.outputs = list(outputs)
        self.constants = dict(constants)
        self.steps = list(steps)
        self._plan_gc()

    def _plan_gc(self):
        last = {}
        for i, st in enumerate(self.steps):
            for s in st.ins:
                last[s] = i
        keep = set(self.inputs) | set(self.outputs) | set(self.constants)
        for st in self.steps:
            st.free = []
        produced = set()
        for st in self.steps:
            produced.update(st.outs)
        for s, i in last.items():
            if s not in keep and s in produced:
                self.steps[i].free.append(s)
        # values never read and not outputs die right after they are produced
        for i, st in enumerate(self.steps):
            for s in st.outs:
                if s not in last and s not in keep:
                    st.free.append(s)


class Executor:
    """Runs a Program.  `run(input_values) -> list of output Vals` (no host conversion)."""

    def __init__(self, program: Program, allow_gc=True):
        self.program = program
        self.allow_gc = allow_gc
        self.vals = [None] * program.n_slots
        for s, arr in program.constants.items():
            self.vals[s] = Val(h=np.asarray(arr))
        self.position_of_error = -1
        self.time_nodes = False
        n = max([len(program.steps)] + [st.origin + 1 for st in program.steps])
        self.call_times = [0.0] * n
        self.call_counts = [0] * n
        self.event_log = None  # when a list: (step index, start event, stop event) per executed step (no syncs)

    def run(self, inputs):
        p = self.program
        vals = self.vals
        for s, x in zip(p.inputs, inputs):
            vals[s] = wrap(x)
        timing = self.time_nodes
        for i, st in enumerate(p.steps):
            try:
                if timing:
                    dev.synchronize()
                    t0 = time.perf_counter()
                if self.event_log is not None:
                    import torch

                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    res = st.impl.run([vals[j] for j in st.ins])
                    e1.record()
                    self.event_log.append((i, e0, e1))
                else:
                    res = st.impl.run([vals[j] for j in st.ins])
                if timing:
                    dev.synchronize()
                    k = st.origin if st.origin >= 0 else i
                    self.call_times[k] += time.perf_counter() - t0
                    self.call_counts[k] += 1
            except Exception:
                self.position_of_error = st.origin if st.origin >= 0 else i
                raise
            for j, r in zip(st.outs, res):
                vals[j] = r
            if self.allow_gc:
                for j in st.free:
                    vals[j] = None
        outs = [vals[s] for s in p.outputs]
        if self.allow_gc:
            for s in p.inputs:
                vals[s] = None
            for s in p.outputs:
                if s not in p.constants:
                    vals[s] = None
        return outs


def outputs_to_host(out_vals, device_outputs=False):
    """Val -> what Function.__call__ hands to the user: NumPy arrays (one sync) or device tensors."""
    res = []
    pending = []
    for v in out_vals:
        if v.h is not None and v.d is None:
            res.append(np.asarray(v.h))
        elif device_outputs:
            res.append(v.d)
        else:
            res.append(None)
            pending.append((len(res) - 1, v))
    for k, v in pending:
        res[k] = dev.to_host(v.d, sync=False)
    if pending or nodes_basic._pending_flags:
        dev.synchronize()
    nodes_basic.check_pending_flags()
    return res
