"""CPU suite: the C-ABI library loads and exports exactly the symbols include/ptk.h declares (no compute calls)."""

import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(REPO, "include", "ptk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ptk_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from pytensor_b200.runtime import lib

    names = _declared()
    assert len(names) >= 35
    L = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ptk.h but not exported by libptk.so"
        assert n in lib.SIGNATURES, f"{n} declared in include/ptk.h but has no ctypes prototype"
    for n in lib.SIGNATURES:
        assert n in names, f"{n} bound in runtime/lib.py but not declared in include/ptk.h"


def test_library_refuses_to_compute_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pytensor_b200.runtime import lib

    L = lib.load_library()
    assert L.ptk_version() >= 100
    assert L.ptk_init(0) != 0
    assert b"no CUDA device" in L.ptk_last_error()
    # every compute entry point fails loudly before ptk_init succeeded
    assert L.ptk_gemv(11, 1, 1, 1.0, None, 1, 1, None, 1, 0.0, None, 1, None) != 0
    assert b"ptk_init" in L.ptk_last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "pytensor_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(root, f)


def test_nvrtc_cross_compiles_sm100a_without_a_gpu():
    from pytensor_b200.runtime import jit

    src = 'extern "C" __global__ void k_cabi_probe(float* x){ x[threadIdx.x] = tanhf(x[threadIdx.x]); }'
    key, cubin = jit.compile_cubin(src)
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 500
