"""The C-ABI library loads and exports every symbol include/hydravox_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from flowmirror_hydravox_b200 import build
    path = build.build()
    lib = ctypes.CDLL(path)
    hdr = open(os.path.join(ROOT, "include", "hydravox_b200.h")).read()
    names = set(re.findall(r"\b(hvx_[a-z0-9_]+)\s*\(", hdr)) - {"hvx_status"}
    assert len(names) >= 10
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    lib.hvx_last_error.restype = ctypes.c_char_p
    assert lib.hvx_version() >= 100


def test_engine_fails_loudly_without_gpu():
    import pytest
    import torch
    from flowmirror_hydravox_b200 import _lib as L
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.HvxError):
        L.Engine()
