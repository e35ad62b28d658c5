"""The C-ABI library loads and exports every symbol include/phase_b200.h declares
(no compute calls: runs without a GPU), and fails loudly without a device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "phase_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from phase_b200 import _capi
    L = _capi.lib()
    names = declared_symbols()
    assert len(names) > 60
    for n in names:
        assert hasattr(L, n), n
        assert n in _capi.SIGNATURES, "missing ctypes signature for " + n
    assert L.phb_version() >= 100


def test_no_cpu_fallback_without_device():
    import ctypes as C
    from phase_b200 import _capi
    L = _capi.lib()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert L.phb_ctx_create(0, C.byref(h)) < 0
    assert b"no CPU fallback" in L.phb_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "phase_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "phase_oracle" not in txt, f
