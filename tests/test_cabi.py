"""The C-ABI shared library loads and exports every entry point include/tpgx.h declares
(no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import tpgx


def declared_functions():
    text = open(os.path.join(ROOT, "include", "tpgx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(tpgx_[a-z0-9_]+)\s*\(", text))
    return sorted(names - {"tpgx_status"})  # `tpgx_status (*callback)(...)` members of tpgx_train_backend


def test_header_and_binding_agree():
    assert declared_functions() == sorted(tpgx.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = tpgx.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert lib.tpgx_abi_version() == 2


def test_struct_sizes_match_header():
    A = tpgx.abi
    assert ctypes.sizeof(A.Line) == 24
    assert ctypes.sizeof(A.Config) == 32
    assert ctypes.sizeof(A.SourceDesc) == 16
    assert ctypes.sizeof(A.Counters) == 48
    assert ctypes.sizeof(A.ClassifRequest) == 32
    assert tpgx.LINE_DTYPE.itemsize == 24


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device tpgx_create must fail, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(tpgx.TpgxError) as e:
        tpgx.Evaluator(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], nb_constants=5)
    assert e.value.status == tpgx.abi.ERR_CUDA
