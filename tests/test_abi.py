"""CPU: the C-ABI library loads and exports every symbol include/facfake.h declares; no compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from fac_fake_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_if_missing():
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.run(["make", "-C", ROOT, "all"], check=True)


def test_header_symbols_exported():
    _build_if_missing()
    hdr = open(os.path.join(ROOT, "include", "facfake.h")).read()
    declared = set(re.findall(r"\b(ff_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("ff_cvit")          # struct tag
    assert len(declared) >= 12
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in facfake.h but not exported"
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)


def test_no_cpu_fallback():
    _build_if_missing()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.ff_cvit_create(C.byref(h), 0, 32, 0)
    assert rc == _lib.FF_ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.ff_last_error(None) or b"CUDA" in lib.ff_last_error(None)
    assert lib.ff_cvit_create(None, 0, 32, 0) == _lib.FF_ERR_BAD_ARG
    assert lib.ff_cvit_create(C.byref(h), 0, 0, 0) == _lib.FF_ERR_BAD_ARG
    assert lib.ff_cvit_create(C.byref(h), 0, 32, 7) == _lib.FF_ERR_BAD_ARG
    # NULL handles are rejected, never dereferenced
    assert lib.ff_cvit_finalize_weights(None) == _lib.FF_ERR_BAD_ARG
    assert lib.ff_cvit_forward(None, None, 0, None, 1, None, None) == _lib.FF_ERR_BAD_ARG
    assert lib.ff_cvit_launch_count(None) == 0
    lib.ff_cvit_destroy(None)


def test_engine_rejects_other_configs_and_cpu():
    _build_if_missing()
    from fac_fake_b200 import CViTEngine, EngineError
    with pytest.raises(ValueError):
        CViTEngine(depth=12)
    with pytest.raises(ValueError):
        CViTEngine(compute_dtype="fp8")
    e = CViTEngine()
    with pytest.raises(EngineError):
        e.to("cpu")
    with pytest.raises(EngineError):
        e.train(True)


def test_product_path_does_not_import_oracle():
    """The product package must never route through oracle/ (SURVEY/tier rule)."""
    pkg = os.path.join(ROOT, "fac_fake_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
