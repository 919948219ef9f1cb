"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/splacu.h declares, and refuses to run (loudly) without a CUDA device -- there is no CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from spla_b200.backend import load_library

    return load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "splacu.h")).read()
    return sorted(set(re.findall(r"SPLACU_API\s+[\w\s\*]+?\b(splacu_\w+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from spla_b200.backend import SYMBOLS

    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/splacu.h but not exported by libsplacu.so"
    assert sorted(SYMBOLS) == names, "spla_b200.backend.SYMBOLS is out of sync with include/splacu.h"


def test_library_is_sm100a():
    from spla_b200 import build

    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.splacu_init(0) == -2  # SPLACU_E_NOT_INIT
    assert b"no CPU fallback" in lib.splacu_last_error()
    h = ctypes.c_void_p()
    assert lib.splacu_workspace_create(ctypes.byref(h)) == -2
    from spla_b200.backend import Backend, SplacuError

    with pytest.raises(SplacuError):
        Backend(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under spla_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spla_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "spla_oracle" not in text and "oracle." not in text and "libspla_ref" not in text, os.path.join(dirpath, f)
