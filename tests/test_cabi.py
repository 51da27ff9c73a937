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


def test_header_is_plain_c_and_links_from_a_c_program(tmp_path):
    """The boundary a non-Python host would bind (cgo / JNI / plain C): include/hydravox_b200.h compiles as C99, a C program links
    against the library, and — with no GPU here — hvx_create fails with a status and a message instead of crashing."""
    import shutil
    import subprocess
    import torch
    from flowmirror_hydravox_b200 import build
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not found")
    lib = build.build()
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "hydravox_b200.h"
int main(void) {
  hvx_config cfg; memset(&cfg, 0, sizeof cfg);
  hvx_engine* e = 0;
  printf("version %d\n", hvx_version());
  hvx_status rc = hvx_create(&e, &cfg);
  printf("create rc=%d err=%s\n", rc, hvx_last_error());
  if (rc == HVX_OK) hvx_destroy(e);
  return 0;
}
''')
    exe = tmp_path / "host"
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe), lib,
                        f"-Wl,-rpath,{os.path.dirname(lib)}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "version " in out.stdout
    if not torch.cuda.is_available():
        assert "create rc=0" not in out.stdout and "err=" in out.stdout and len(out.stdout.split("err=")[1].strip()) > 0
