"""CPU suite: a fixed batch of random graphs through the lowering, each lowered program interpreted by the NumPy port oracle
against the reference C linker (tests/lowering_fuzz.py).  (The graph with which this search exposed the region-ordering bug
is spelled out in tests/test_rowfuse_cpu.py.)"""

import pytest

from lowering_fuzz import check_seed


@pytest.mark.parametrize("first", [0, 20, 40])
def test_random_graphs_lower_to_programs_that_agree_with_the_c_linker(first):
    results = [check_seed(s) for s in range(first, first + 20)]
    assert results.count("ok") >= 12, results


def test_generated_scalar_expressions_follow_c_conversions_like_the_reference():
    """tests/codegen_dtype_fuzz.py over a fixed batch: includes seed 1388, where Maximum(int8, uint32) -> int64 used to keep the
    operand the UNSIGNED comparison selects as an unsigned value, while the reference's expression (one form with a
    nan("") arm for every dtype) routes it through a double and keeps the sign."""
    from codegen_dtype_fuzz import check_seed as check_codegen

    results = [check_codegen(s) for s in list(range(1380, 1392)) + list(range(0, 24))]
    assert results.count("ok") >= 6, results


def test_stream_plans_of_random_programs_order_every_dependency():
    """Program.plan_streams (the capture-time multi-stream schedule) on the random graphs: every read of a produced slot and
    every in-place write after a read of the same slot is ordered — by program order on one stream or by a chain of
    event edges — checked with an independent reachability walk."""
    import numpy as np

    from helpers import pytensor
    from lowering_fuzz import build

    multi = 0
    for seed in range(500, 560):
        rng = np.random.default_rng(seed)
        dtype = "float32" if seed % 2 else "float64"
        pytensor.config.floatX = dtype
        try:
            ins, outs, _ = build(rng, dtype)
            prog = pytensor.function(ins, outs, mode="CUDA", on_unused_input="ignore").vm.executor.program
        except Exception:  # noqa: BLE001  (ill-shaped random graph)
            continue
        prog.plan_streams()
        before, last_on = [set() for _ in prog.steps], {}
        for i in range(len(prog.steps)):
            preds = set(prog.deps[i]) | ({last_on[prog.stream_of[i]]} if prog.stream_of[i] in last_on else set())
            for p in preds:
                before[i] |= {p} | before[p]
            last_on[prog.stream_of[i]] = i
        producer, readers = {}, {}
        for i, st in enumerate(prog.steps):
            for s in st.ins:
                assert s not in producer or producer[s] in before[i], (seed, "read before its producer is ordered", i)
            for k in (getattr(st.impl, "destroy", None) or {}).values():
                if k < len(st.ins):
                    assert all(r in before[i] for r in readers.get(st.ins[k], []) if r != i), (seed, "in-place write races a reader", i)
            for s in st.ins:
                readers.setdefault(s, []).append(i)
            for s in st.outs:
                producer[s] = i
        multi += prog.n_streams > 1
    assert multi >= 20
