"""Row-fused regions on the device (codegen/rowfuse.py): the batched hierarchical logp+grad graph (BASELINE.json
configs[4]) as ONE kernel vs the reference C linker, the unfused fallback, per-row outputs, out-of-bounds indices."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt
from pytensor_b200 import workloads as W
from pytensor_b200.link.cuda import cuda_mode
from pytensor_b200.vm.nodes_rowfuse import RowRegionNode

pytestmark = pytest.mark.gpu


def _region(f):
    nodes = [st.impl for st in f.vm.executor.program.steps if isinstance(st.impl, RowRegionNode)]
    assert len(nodes) == 1
    return nodes[0]


@pytest.mark.parametrize("dtype,B,n,J,K", [("float32", 4096, 1024, 64, 8), ("float32", 777, 100, 13, 5),
                                           ("float64", 512, 256, 32, 8)])
def test_logp_grad_region_matches_cvm(gpu, dtype, B, n, J, K):
    pytensor.config.floatX = dtype
    ins, outs, mk, meta = W.cfg5_logp_grad(B=B, n=n, J=J, K=K, dtype=dtype, packed=False)
    args = mk(seed=31)
    f_ref = pytensor.function(ins, outs, mode="CVM")
    exp = f_ref(*args)
    f = pytensor.function(ins, outs, mode="CUDA")
    node = _region(f)
    for call in range(4):  # eager, capture, replay, replay
        got = f(*args)
        for g, e in zip(got, exp):
            scale = max(1.0, float(np.max(np.abs(e))))
            np.testing.assert_allclose(g, e, rtol=1e-5, atol=1e-5 * scale)
    assert node.fused_calls >= 2 and node.unfused_calls == 0, node.last_reason
    pytensor.config.floatX = "float32"


def test_region_falls_back_to_its_steps(gpu, monkeypatch):
    pytensor.config.floatX = "float32"
    ins, outs, mk, meta = W.cfg5_logp_grad(B=300, n=128, J=16, K=4, packed=True)
    args = mk(seed=32)
    f = pytensor.function(ins, outs, mode=cuda_mode(use_graph=False))
    fused = f(*args)[0]
    node = _region(f)
    assert node.fused_calls == 1
    monkeypatch.setenv("PTK_ROWFUSE", "0")
    unfused = f(*args)[0]
    assert node.unfused_calls == 1
    np.testing.assert_allclose(fused, unfused, rtol=2e-5, atol=2e-3)
    monkeypatch.delenv("PTK_ROWFUSE")
    # too few rows for the fused kernel: the node runs its steps, results still match the C linker
    small = mk(seed=33, B_local=8)
    exp = pytensor.function(ins, outs, mode="CVM")(*small)[0]
    np.testing.assert_allclose(f(*small)[0], exp, rtol=1e-5, atol=1e-4)
    assert node.unfused_calls == 2 and "rows" in node.last_reason


def test_region_with_per_row_outputs_and_negative_indices(gpu):
    pytensor.config.floatX = "float32"
    w, s = pt.fmatrix("w"), pt.fvector("s")
    idx = pt.lvector("idx")
    o1 = pt.exp(w[:, idx] * s[:, None])
    o2 = o1.sum(axis=1)
    rng = np.random.default_rng(5)
    args = [rng.standard_normal((3000, 9)).astype("float32"), rng.standard_normal(3000).astype("float32") * 0.3,
            rng.integers(-9, 9, size=333).astype("int64")]
    f, _ = compare_cuda_and_cvm([w, s, idx], [o1, o2], args)
    assert _region(f).fused_calls >= 1


def test_region_raises_index_error(gpu):
    pytensor.config.floatX = "float32"
    ins, outs, mk, meta = W.cfg5_logp_grad(B=256, n=64, J=8, K=4, packed=True)
    args = mk(seed=34)
    args[6] = args[6].copy()
    args[6][5] = 8
    f = pytensor.function(ins, outs, mode="CUDA")
    with pytest.raises(IndexError):
        f(*args)
    args[6][5] = -8   # wraps like NumPy
    f(*args)
    # device outputs never synchronise: the error surfaces at the next call / on check_errors()
    fd = pytensor.function(ins, outs, mode=cuda_mode(device_outputs=True))
    args[6][5] = 9
    fd(*args)
    with pytest.raises(IndexError):
        fd.vm.check_errors()
