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
