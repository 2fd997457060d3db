"""CPU suite: a fixed batch of random graphs through the lowering, each lowered program interpreted by the NumPy port oracle
against the reference C linker (tests/lowering_fuzz.py).  (The graph with which this search exposed the region-ordering bug
is spelled out in tests/test_rowfuse_cpu.py.)"""

import pytest

from lowering_fuzz import check_seed


@pytest.mark.parametrize("first", [0, 20, 40])
def test_random_graphs_lower_to_programs_that_agree_with_the_c_linker(first):
    results = [check_seed(s) for s in range(first, first + 20)]
    assert results.count("ok") >= 12, results
