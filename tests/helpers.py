"""Backend test idiom of the reference (tests/link/pytorch/test_basic.py:43-87 `compare_pytorch_and_py`): compile the
same graph with the backend mode and with the oracle mode, assert closeness.  Here the oracle is the reference's C
linker (`mode="CVM"`), as BASELINE.json's north_star demands."""

import os

import numpy as np

from oracle import cvm

pytensor = cvm.configure()
import pytensor_b200  # noqa: E402,F401  (registers mode="CUDA")


def tol_for(dtype):
    # the reference's own notion of "close" (pytensor/tensor/math.py:92-103): fp32 1e-5, fp64 rtol 1e-5
    return dict(rtol=1e-5, atol=1e-5) if np.dtype(dtype) == np.float32 else dict(rtol=1e-5, atol=1e-8)


def compare_cuda_and_cvm(inputs, outputs, test_inputs, mode="CUDA", rtol=None, atol=None, exact=False, cvm_kwargs=None,
                         atol_scale=None):
    """`atol_scale`: absolute tolerance as a fraction of each expected output's largest magnitude (the natural unit of a
    dot product's rounding error: 1e-5 of the output scale is the north star's bar for fp32 Gemm)."""
    single = not isinstance(outputs, list | tuple)
    outs = [outputs] if single else list(outputs)
    f_cuda = pytensor.function(inputs, outs, mode=mode)
    f_ref = pytensor.function(inputs, outs, mode="CVM", **(cvm_kwargs or {}))
    if os.environ.get("PTK_DRY") == "1":  # developer dry run without a GPU: lowering + reference only
        from pytensor_b200.precompile import trace_function

        exp = f_ref(*[np.array(x, copy=True) for x in test_inputs])
        trace_function(f_cuda, [np.array(x, copy=True) for x in test_inputs])  # launch logic + NVRTC, no device
        # ... and the LOWERED program (after this backend's fusion passes and peepholes) interpreted by the NumPy port
        # oracle against the C linker: checks the lowering itself, not the kernels
        from oracle import numpy_port

        prog = f_cuda.vm.executor.program
        if len(prog.inputs) == len(test_inputs):
            try:
                got = numpy_port.evaluate_program(prog, [np.array(x, copy=True) for x in test_inputs])
            except NotImplementedError:
                got = None
            if got is not None:
                _assert_close(got, exp, rtol, atol, exact, atol_scale, port=True)
        return f_cuda, None
    got = f_cuda(*[np.array(x, copy=True) for x in test_inputs])
    exp = f_ref(*[np.array(x, copy=True) for x in test_inputs])
    _assert_close(got, exp, rtol, atol, exact, atol_scale)
    return f_cuda, got


def _assert_close(got, exp, rtol, atol, exact, atol_scale, port=False):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        g, e = np.asarray(g), np.asarray(e)
        assert g.dtype == e.dtype, (g.dtype, e.dtype)
        assert g.shape == e.shape, (g.shape, e.shape)
        if (exact and not (port and e.dtype.kind == "f")) or e.dtype.kind in "biu":
            np.testing.assert_array_equal(g, e)
        else:
            t = tol_for(e.dtype)
            a = atol if atol is not None else t["atol"]
            if atol_scale is not None and e.size:
                a = atol_scale * float(np.max(np.abs(e[np.isfinite(e)]))) if np.isfinite(e).any() else a
            np.testing.assert_allclose(g, e, rtol=rtol if rtol is not None else t["rtol"], atol=a, equal_nan=True)
