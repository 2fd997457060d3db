"""VM / linker contract on the device (after tests/link/test_vm.py): CUDA-graph replay equals eager execution, new
shapes re-specialise, updates, profiling hooks, the metric graph end-to-end."""

import numpy as np
import pytest

from helpers import compare_cuda_and_cvm, pytensor

import pytensor.tensor as pt
from pytensor_b200.link.cuda import cuda_mode

pytestmark = pytest.mark.gpu


def test_graph_replay_matches_eager_and_respecialises(gpu):
    rng = np.random.default_rng(61)
    x, y = pt.dmatrix("x"), pt.dvector("y")
    out = [pt.tanh(pt.dot(x, x.T)).sum(axis=0) + y, (x * y[:, None]).max()]
    f_g = pytensor.function([x, y], out, mode=cuda_mode(use_graph=True))
    f_e = pytensor.function([x, y], out, mode=cuda_mode(use_graph=False))
    for shape in [(12, 5), (12, 5), (12, 5), (12, 5), (7, 9), (7, 9), (7, 9), (12, 5)]:
        xv, yv = rng.standard_normal(shape), rng.standard_normal(shape[0])
        a, b = f_g(xv, yv), f_e(xv, yv)
        for p, q in zip(a, b):
            np.testing.assert_allclose(p, q, rtol=1e-12)
    assert f_g.vm.executor.last_from_graph
    assert any(e.stage == 2 for e in f_g.vm.executor._graphs.values())


def test_graph_replay_with_scalar_inputs_and_shared_updates(gpu):
    w = pytensor.shared(np.zeros(5), name="w")
    x = pt.dvector("x")
    lr = pt.dscalar("lr")
    f = pytensor.function([x, lr], (w * x).sum(), updates={w: w + lr * x}, mode="CUDA")
    xv = np.arange(5.0)
    expect_w = np.zeros(5)
    for k in range(6):
        lrv = 0.1 * (k % 2 + 1)  # two alternating scalar values -> two signatures
        got = f(xv, lrv)
        np.testing.assert_allclose(got, (expect_w * xv).sum(), rtol=1e-12)
        expect_w = expect_w + lrv * xv
        np.testing.assert_allclose(w.get_value(), expect_w, rtol=1e-12)


def test_outputs_are_fresh_objects(gpu):
    x = pt.dvector("x")
    f = pytensor.function([x], x * 2, mode="CUDA")
    outs = [f(np.full(4, float(k))) for k in range(5)]
    for k, o in enumerate(outs):
        np.testing.assert_array_equal(o, np.full(4, 2.0 * k))


def test_profile_hooks(gpu):
    x = pt.dmatrix("x")
    f = pytensor.function([x], pt.exp(x).sum(axis=1), mode="CUDA", profile=True)
    f(np.ones((4, 4)))
    f(np.ones((4, 4)))
    assert f.profile.fct_callcount >= 1
    assert sum(f.profile.apply_callcount.values()) >= 1


def test_metric_graph_small(gpu):
    # the 256-node-class graph of BASELINE.json's metric at reduced depth: Dot22 + Elemwise layers, a Scan, a Sum
    from pytensor_b200 import workloads as W

    pytensor.config.floatX = "float32"
    ins, outs, make_args, _ = W.metric_graph(n=32, layers=12, scan_steps=8)
    f, _ = compare_cuda_and_cvm(ins, outs, make_args(), rtol=1e-4, atol=1e-5)
    for _ in range(3):
        f(*make_args())
    assert f.vm.executor.last_from_graph


def test_cfg5_logp_grad_small(gpu):
    from pytensor_b200 import workloads as W

    pytensor.config.floatX = "float64"
    try:
        ins, outs, make_args, _ = W.cfg5_logp_grad(B=300, n=128, J=16, K=4, dtype="float64")
        compare_cuda_and_cvm(ins, outs, make_args(), rtol=1e-8, atol=1e-8)
    finally:
        pytensor.config.floatX = "float32"


def test_op_from_graph_and_batched_dot(gpu):
    from pytensor.compile.builders import OpFromGraph

    rng = np.random.default_rng(62)
    x, y = pt.dmatrix("x"), pt.dmatrix("y")
    ofg = OpFromGraph([x, y], [pt.tanh(pt.dot(x, y)) + x.sum(), pt.exp(-x)], inline=False)
    a, b = pt.dmatrix("a"), pt.dmatrix("b")
    o1, o2 = ofg(a, b)
    compare_cuda_and_cvm([a, b], [o1 * 2, o2.sum(axis=0)], [rng.standard_normal((5, 5)), rng.standard_normal((5, 5))],
                         rtol=1e-9, atol=1e-10)
    A3, B3 = pt.dtensor3("A3"), pt.dtensor3("B3")
    compare_cuda_and_cvm([A3, B3], [pt.matmul(A3, B3)], [rng.standard_normal((4, 6, 7)), rng.standard_normal((4, 7, 3))],
                         rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("shape,dtype", [((1030, 4099), "float32"), ((3_000_001,), "float64"), ((37, 129, 515), "float32"),
                                         ((2048, 2048), "int64")])
def test_chunked_host_pipeline_matches_cvm(gpu, shape, dtype):
    # large host-resident elementwise call: the executor pipelines H2D | kernel | D2H chunk by chunk along axis 0
    # (Executor._run_chunked); results must equal the one-shot path and the C linker, including ragged last chunks
    rng = np.random.default_rng(67)
    T = pt.TensorType(dtype, shape=(None,) * len(shape))
    x, y, z = T("x"), T("y"), T("z")
    if dtype == "int64":
        outs = [x * y - z, pt.maximum(x, y) + (z & 7)]
        vals = [rng.integers(-1000, 1000, size=shape).astype(dtype) for _ in range(3)]
    else:
        outs = [pt.tanh(x * y) + pt.exp(-z * z), x - y * z]
        vals = [rng.standard_normal(shape).astype(dtype) for _ in range(3)]
    f = pytensor.function([x, y, z], outs, mode="CUDA")
    f_ref = pytensor.function([x, y, z], outs, mode="CVM")
    for _ in range(3):  # eager + arena measurement, per-chunk graph capture, per-chunk graph replay
        got = f(*vals)
    assert f.vm.executor.chunked_calls == 3
    assert len(f.vm.executor._chunk_plan["bounds"]) >= 2
    f.vm.executor.STREAM_MIN_BYTES = 1 << 60  # same function, one-shot path
    whole = f(*vals)
    assert f.vm.executor.chunked_calls == 3
    for g, w, r in zip(got, whole, f_ref(*vals)):
        assert g.shape == r.shape and g.dtype == r.dtype
        np.testing.assert_array_equal(g, w)
        if dtype == "int64":
            np.testing.assert_array_equal(g, r)
        else:
            np.testing.assert_allclose(g, r, rtol=1e-5 if dtype == "float32" else 1e-12, atol=1e-6)


def test_chunked_pipeline_not_taken_for_broadcast_or_small(gpu):
    x, y = pt.fmatrix("x"), pt.fvector("y")
    f = pytensor.function([x, y], x * y + 1, mode="CUDA")
    xv = np.ones((4096, 4096), "float32")
    r = f(xv, np.arange(4096, dtype="float32"))
    assert f.vm.executor.chunked_calls == 0
    np.testing.assert_array_equal(r[5], np.arange(4096, dtype="float32") + 1)
    g = pytensor.function([x], pt.exp(x), mode="CUDA")
    g(np.ones((64, 64), "float32"))
    assert g.vm.executor.chunked_calls == 0


def test_chunked_pipeline_with_row_reduction_cfg2(gpu):
    # BASELINE.json configs[1] at reduced size: fused Elemwise -> row sum keeps axis 0, so host calls are pipelined too;
    # chunking must not change a single bit (rows are reduced independently)
    from pytensor_b200 import workloads as W

    pytensor.config.floatX = "float32"
    ins, outs, make_args, _ = W.cfg2_fused_elemwise(2048)
    f = pytensor.function(ins, outs, mode="CUDA")
    args = make_args(71)
    for _ in range(4):
        got = f(*args)
    assert f.vm.executor.chunked_calls == 4
    assert f.vm.executor._chunk_plan["sub"].last_from_graph
    f.vm.executor.STREAM_MIN_BYTES = 1 << 60
    whole = f(*args)
    ref = pytensor.function(ins, outs, mode="CVM")(*args)
    for g, w, r in zip(got, whole, ref):
        np.testing.assert_array_equal(g, w)
        np.testing.assert_allclose(g, r, rtol=1e-5, atol=1e-6)


def test_opfromgraph_does_not_clobber_its_inputs(gpu):
    """An OpFromGraph declares no destroy_map: its inner graph must not write into the outer value it was handed
    (reference: destructive_rewrite_ofg_inner_graph protects the inner inputs, compile/rewriting.py:141-152)."""
    from pytensor.compile.builders import OpFromGraph

    xi = pt.fmatrix("xi")
    ofg = OpFromGraph([xi], [pt.exp(-xi)])
    x = pt.fmatrix("x")
    y = x * np.float32(2)
    w = y + ofg(y)  # y is read again AFTER the OpFromGraph node
    xv = np.random.default_rng(71).standard_normal((33, 65)).astype("float32")
    compare_cuda_and_cvm([x], [w], [xv])
    # an inner output that is a view of an inner input must come back as a fresh buffer
    ofg_t = OpFromGraph([xi], [xi.T])
    compare_cuda_and_cvm([x], [ofg_t(y) + np.float32(1), y], [xv])


def test_opfromgraph_leaves_device_shared_variables_intact(gpu):
    from pytensor.compile.builders import OpFromGraph
    import pytensor_b200

    xi = pt.fvector("xi")
    ofg = OpFromGraph([xi], [pt.tanh(xi) * np.float32(3)])
    v0 = np.linspace(-1, 1, 257).astype("float32")
    s = pytensor_b200.shared(v0.copy(), name="s")
    f = pytensor.function([], ofg(s) + s, mode="CUDA")
    for _ in range(4):  # eager, capture, replay, replay
        np.testing.assert_allclose(f(), np.tanh(v0) * 3 + v0, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(s.get_value(), v0)
