"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/mjxb.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers
from mujoco_mjx_lab_b200 import _abi, _lib, modelc


def _declared_symbols():
    src = open(os.path.join(helpers.ROOT, "include", "mjxb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mjxb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_listed_and_exported():
    declared = _declared_symbols()
    assert sorted(_lib.SYMBOLS) == declared
    _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"libmjxb.so does not export {name}"
    assert L.mjxb_abi_version() == 1


def test_struct_sizes_agree():
    L = _lib.lib()
    assert L.mjxb_blob_sizeof() == modelc.BLOB_DTYPE.itemsize
    assert L.mjxb_env_config_sizeof() == ctypes.sizeof(_abi.EnvConfigC)


def test_model_create_errors_without_compute(model):
    """Argument / blob validation happens on the host; without a GPU the create call reports ENOGPU (no CPU fallback)."""
    import torch
    L = _lib.lib()
    blob = modelc.pack_blob(model)
    h = ctypes.c_void_p()
    assert L.mjxb_model_create(None, blob.nbytes, None, 0, ctypes.byref(h)) == -1
    assert L.mjxb_model_create(blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes - 4, None, 0, ctypes.byref(h)) == -2
    bad = blob.copy(); bad["magic"] = 7
    assert L.mjxb_model_create(bad.ctypes.data_as(ctypes.c_void_p), bad.nbytes, None, 0, ctypes.byref(h)) == -2
    if not torch.cuda.is_available():
        assert L.mjxb_model_create(blob.ctypes.data_as(ctypes.c_void_p), blob.nbytes, None, 0, ctypes.byref(h)) == -4
        from mujoco_mjx_lab_b200 import mjx
        with pytest.raises(_lib.MjxbError):
            mjx.put_model(model)
    assert b"no CPU fallback" in L.mjxb_strerror(-4)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, load or include it."""
    pkg = os.path.join(helpers.ROOT, "mujoco_mjx_lab_b200")
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle\.hpp|oracle/|#include\s+\"oracle)")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), f"{f} references the oracle"
