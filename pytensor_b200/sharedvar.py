"""Device-resident shared variables.

`pytensor.shared(ndarray)` keeps its value as a NumPy array (`TensorSharedVariable`, pytensor/tensor/sharedvar.py:32-94;
`SharedVariable.get_value/set_value`, pytensor/compile/sharedvalue.py:97-136), so every call of a CUDA function would
upload it again and every `updates=` write would come back over PCIe (`Function.__call__` stores the update output
straight into the container's cell, pytensor/compile/executor.py:712-716).  `pytensor_b200.shared(value)` returns a
`CudaSharedVariable`:

  * symbolically it is an ordinary `TensorSharedVariable` of a plain `TensorType` — every rewrite, `grad`, `updates=`
    and `givens=` treats it exactly like the reference's (a `TensorType` subclass would not even count as dense:
    `DenseTypeMeta.__instancecheck__`, pytensor/tensor/type.py:625-629);
  * its container's cell may hold either a NumPy array (after construction / `set_value(ndarray)`) or a torch CUDA
    tensor.  The CUDA VM promotes a NumPy value to HBM on the first call that needs it and leaves the device tensor in
    the cell; updates are written into that same buffer on the device, so the address stays stable and the captured
    CUDA graph keeps replaying (`CudaVM.__call__`, link/cuda/linker.py);
  * `get_value()` copies device -> host on demand, `get_value(borrow=True, return_internal_type=True)` hands out the
    device tensor itself.

Such a variable can only be an input of functions compiled with a CUDA mode: a C/Python thunk would find a device
tensor in its storage cell.  That is the same restriction the value's owner accepts by asking for device residency.
"""

from __future__ import annotations

import numpy as np

from pytensor_b200._host import ensure_pytensor

ensure_pytensor()

from pytensor.tensor.sharedvar import TensorSharedVariable  # noqa: E402
from pytensor.tensor.type import TensorType  # noqa: E402


def _is_dev(v) -> bool:
    return hasattr(v, "is_cuda") and hasattr(v, "data_ptr")


class CudaSharedVariable(TensorSharedVariable):
    """`TensorSharedVariable` whose value lives in HBM between calls of CUDA-mode functions."""

    def _check_dev(self, t):
        from pytensor_b200.runtime import device as dev

        dt = dev.TORCH_TO_NP.get(t.dtype)
        if dt != self.type.dtype or t.dim() != self.type.ndim:
            raise TypeError(f"{self}: device value of dtype {dt}, ndim {t.dim()} for type {self.type}")
        for have, want in zip(t.shape, self.type.shape):
            if want is not None and int(have) != want:
                raise TypeError(f"{self}: device value of shape {tuple(t.shape)} for type {self.type}")

    def get_value(self, borrow=False, return_internal_type=False):
        v = self.container.storage[0]
        if _is_dev(v):
            from pytensor_b200.runtime import device as dev

            if return_internal_type:
                return v if borrow else dev.clone(v)
            return dev.to_host(v)  # a fresh host array either way (synchronises the VM stream)
        return super().get_value(borrow=borrow, return_internal_type=return_internal_type)

    def set_value(self, new_value, borrow=False):
        if _is_dev(new_value):
            from pytensor_b200.runtime import device as dev

            self._check_dev(new_value)
            self.container.storage[0] = new_value if borrow else dev.clone(new_value)
            return
        # host value: validated by TensorType.filter through the container; promoted by the next CUDA call
        super().set_value(new_value, borrow=borrow)

    def zero(self, borrow=False):
        v = self.container.storage[0]
        if _is_dev(v):
            from pytensor_b200.runtime import device as dev

            z = np.zeros((), dtype=self.type.dtype)
            src = dev.to_device(z).as_strided(tuple(v.shape), (0,) * v.dim())
            if borrow:
                dev.copy_strided(v, src)
            else:
                out = dev.empty(tuple(v.shape), self.type.dtype)
                dev.copy_strided(out, src)
                self.container.storage[0] = out
            return
        super().zero(borrow=borrow)

    @property
    def on_device(self) -> bool:
        return _is_dev(self.container.storage[0])


def shared(value, name=None, strict=False, allow_downcast=None, borrow=False, shape=None) -> CudaSharedVariable:
    """Device-resident counterpart of `pytensor.shared` for array values (tensor_constructor,
    pytensor/tensor/sharedvar.py:53-94: all dims resizable unless `shape` says otherwise)."""
    if _is_dev(value):
        from pytensor_b200.runtime import device as dev

        dtype = dev.TORCH_TO_NP[value.dtype]
        if shape is None:
            shape = (None,) * value.dim()
        var = CudaSharedVariable(type=TensorType(dtype, shape=shape),
                                 value=np.zeros([1 if s is None else s for s in shape], dtype=dtype), strict=strict,
                                 allow_downcast=allow_downcast, name=name)
        var.set_value(value, borrow=borrow)
        return var
    value = np.asarray(value)
    if isinstance(value, np.ma.MaskedArray):
        raise NotImplementedError("MaskedArrays are not supported")
    if shape is None:
        shape = (None,) * value.ndim
    return CudaSharedVariable(type=TensorType(value.dtype, shape=shape), value=np.array(value, copy=(not borrow)),
                              strict=strict, allow_downcast=allow_downcast, name=name)
