"""Device-resident shared variables: `pytensor_b200.shared` keeps parameters in HBM across calls and
writes `updates=` on the device; results must equal the reference C linker driving ordinary `pytensor.shared` values
(the reference's shared-variable contract: tests/compile/test_shared.py, tests/compile/function/test_pfunc.py)."""

import numpy as np
import pytest

from helpers import pytensor

import pytensor.tensor as pt
import pytensor_b200

pytestmark = pytest.mark.gpu


def _sgd_functions(shared_ctor, mode):
    rng = np.random.default_rng(81)
    W = shared_ctor(rng.standard_normal((64, 32)).astype("float32") / 8, name="W")
    b = shared_ctor(np.zeros(32, "float32"), name="b")
    x, t = pt.fmatrix("x"), pt.fmatrix("t")
    y = pt.tanh(pt.dot(x, W) + b)
    loss = ((y - t) ** 2).mean()
    gW, gb = pytensor.grad(loss, [W, b])
    lr = np.float32(0.5)
    step = pytensor.function([x, t], loss, updates={W: W - lr * gW, b: b - lr * gb}, mode=mode)
    predict = pytensor.function([x], y, mode=mode)
    return W, b, step, predict


def test_sgd_with_device_resident_parameters_matches_cvm(gpu):
    pytensor.config.floatX = "float32"
    Wd, bd, step_d, pred_d = _sgd_functions(pytensor_b200.shared, "CUDA")
    Wr, br, step_r, pred_r = _sgd_functions(pytensor.shared, "CVM")
    rng = np.random.default_rng(82)
    ptrs = set()
    for i in range(8):
        xv = rng.standard_normal((128, 64)).astype("float32")
        tv = np.tanh(xv[:, :32])
        ld, lr_ = step_d(xv, tv), step_r(xv, tv)
        np.testing.assert_allclose(ld, lr_, rtol=1e-5, atol=1e-6)
        assert Wd.on_device and bd.on_device
        ptrs.add((Wd.container.storage[0].data_ptr(), bd.container.storage[0].data_ptr()))
    assert len(ptrs) == 1, "updates must land in the same device buffers (stable addresses)"
    assert step_d.vm.executor.last_from_graph, "stable addresses -> the captured CUDA graph replays"
    np.testing.assert_allclose(Wd.get_value(), Wr.get_value(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(bd.get_value(), br.get_value(), rtol=1e-5, atol=1e-6)
    # a second function sees the values the first one left on the device
    xv = rng.standard_normal((16, 64)).astype("float32")
    np.testing.assert_allclose(pred_d(xv), pred_r(xv), rtol=1e-5, atol=1e-6)


def test_set_value_get_value_zero_and_internal_type(gpu):
    import torch

    v = pytensor_b200.shared(np.arange(6, dtype="float64").reshape(2, 3), name="v")
    x = pt.dmatrix("x")
    f = pytensor.function([x], x + v, mode="CUDA")
    np.testing.assert_array_equal(f(np.ones((2, 3))), np.arange(6.0).reshape(2, 3) + 1)
    assert v.on_device
    p0 = v.container.storage[0].data_ptr()
    got = v.get_value()
    assert isinstance(got, np.ndarray) and got.dtype == np.float64
    np.testing.assert_array_equal(got, np.arange(6.0).reshape(2, 3))
    # host set_value of the same layout: uploaded into the same buffer at the next call
    v.set_value(np.full((2, 3), 5.0))
    assert not v.on_device
    np.testing.assert_array_equal(f(np.ones((2, 3))), np.full((2, 3), 6.0))
    assert v.on_device and v.container.storage[0].data_ptr() == p0
    # new shape: new buffer, still correct
    v.set_value(np.full((4, 3), 2.0))
    np.testing.assert_array_equal(f(np.ones((4, 3))), np.full((4, 3), 3.0))
    # device-side accessors
    t = v.get_value(borrow=True, return_internal_type=True)
    assert isinstance(t, torch.Tensor) and t.is_cuda and tuple(t.shape) == (4, 3)
    v.zero(borrow=True)
    np.testing.assert_array_equal(v.get_value(), np.zeros((4, 3)))
    v.set_value(torch.full((2, 3), 7.0, dtype=torch.float64, device="cuda"))
    np.testing.assert_array_equal(f(np.zeros((2, 3))), np.full((2, 3), 7.0))
    with pytest.raises(TypeError):
        v.set_value(torch.zeros((2, 3), dtype=torch.float32, device="cuda"))


def test_updates_that_change_shape_or_alias_and_scalar_counters(gpu):
    # growing vector (new buffer every call), a reversed VIEW of the old value (overlapping memory), an int64 counter
    def build(shared_ctor, mode):
        v = shared_ctor(np.arange(3, dtype="float64"), name="v")
        w = shared_ctor(np.arange(5, dtype="float64"), name="w")
        c = shared_ctor(np.asarray(0, dtype="int64"), name="c")
        f = pytensor.function([], [v.sum(), c * 2], updates={v: pt.concatenate([v, v * 2]), w: w[::-1], c: c + 1},
                              mode=mode)
        return v, w, c, f

    vd, wd, cd, fd = build(pytensor_b200.shared, "CUDA")
    vr, wr, cr, fr = build(pytensor.shared, "CVM")
    for _ in range(4):
        a, b = fd(), fr()
        np.testing.assert_allclose(a[0], b[0])
        np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_allclose(vd.get_value(), vr.get_value())
    np.testing.assert_array_equal(wd.get_value(), wr.get_value())
    assert int(cd.get_value()) == int(cr.get_value()) == 4
    assert vd.on_device and wd.on_device and cd.on_device


def test_update_only_function_is_asynchronous_and_correct(gpu):
    # no returned outputs: nothing crosses PCIe, nothing synchronises; the value is right when finally read
    acc = pytensor_b200.shared(np.zeros((256, 256), "float32"), name="acc")
    x = pt.fmatrix("x")
    f = pytensor.function([x], [], updates={acc: acc + pt.dot(x, x.T)}, mode="CUDA")
    rng = np.random.default_rng(83)
    xs = [rng.standard_normal((256, 64)).astype("float32") for _ in range(5)]
    for xv in xs:
        f(xv)
    ref = sum(xv.astype(np.float64) @ xv.T.astype(np.float64) for xv in xs)
    np.testing.assert_allclose(acc.get_value(), ref, rtol=1e-5, atol=1e-4)
