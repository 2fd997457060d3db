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
