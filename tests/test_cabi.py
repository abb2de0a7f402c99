"""The drop-in boundary: the shared library loads, exports every symbol include/deepimpute_b200.h declares, the
ctypes binding covers all of them, and without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _has_gpu
from deepimpute_b200 import _lib

HEADER = os.path.join(ROOT, "include", "deepimpute_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(di_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES)
    assert lib.di_version() >= 100


def test_exported_symbols_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for name in declared_symbols():
        assert name in exported                  # unmangled: extern "C"


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "deepimpute_b200.h"\nint main(void){ di_config c; (void)c; return DI_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_config_struct_layout_matches_header():
    # di_config: 4 x int32, 5 x float, uint64, 2 x int32 -> the uint64 forces 8-byte alignment
    assert C.sizeof(_lib.DiConfig) == 56
    assert _lib.DiConfig.seed.offset == 40 and _lib.DiConfig.math_mode.offset == 48


def test_bad_arguments_are_rejected_before_touching_the_device():
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.DiConfig(0, 256, 512, 64, 1e-4, 0.9, 0.999, 1e-7, 0.2, 1, 0, 0)
    n_pred = np.array([10], dtype=np.int32)
    assert lib.di_create(C.byref(h), C.byref(cfg), _lib.i32(n_pred)) == 1          # DI_ERR_ARG: S = 0
    assert b"invalid" in lib.di_last_error(None)
    assert lib.di_predict(None, None, 0, None) == 1
    assert lib.di_launch_count(None) == 0
    # the fused ends follow the same convention: a missing handle is an argument error
    raw = np.zeros((2, 2), dtype=np.float32)
    assert lib.di_upload_counts(None, C.c_void_p(raw.ctypes.data), 0, 2, 2) == 1
    assert lib.di_impute(None, 1, None, 0, None, 0, 1, C.c_void_p(raw.ctypes.data)) == 1
    # stand-alone predictor selection validates its arguments before it looks for a device
    top = np.zeros(5, dtype=np.int32)
    rc = lib.di_corr_topk(0, None, 0, 0, None, 0, None, 0, 0, 5, _lib.i32(top), None, None)
    assert rc != 0 and lib.di_corr_last_error()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback():
    from deepimpute_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        Engine([16, 16], hidden=8, sub_outputdim=8)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdeepimpute_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under deepimpute_b200/ may import, call or mention it as code."""
    pkg = os.path.join(ROOT, "deepimpute_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, name)
