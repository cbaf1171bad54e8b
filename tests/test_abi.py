"""The C-ABI library loads on a CPU-only box, exports every symbol include/wvb.h declares, and refuses
to decode without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from wavpackdecoder_b200 import build, _native
    build.build()
    return _native.load()


def test_exports_every_declared_symbol(lib):
    from wavpackdecoder_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "wvb.h")).read()
    declared = set(re.findall(r"\b(wvb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_sizes_match_header(lib):
    from wavpackdecoder_b200 import _native as N
    assert C.sizeof(N.BlockDesc) == 160
    assert C.sizeof(N.BlockResult) == 16
    assert lib.wvb_abi_version() == 3
    N.check_layout(lib)  # every field offset and size of the ctypes mirrors against the compiled C structs


def test_no_cpu_fallback(lib):
    if lib.wvb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    from wavpackdecoder_b200 import _native as N
    assert lib.wvb_batch_create(0, C.byref(h)) == N.E_NO_DEVICE
    assert b"no CPU decode path" in lib.wvb_last_error()


def test_product_does_not_touch_oracle():
    """Nothing under the package (or the C ABI sources) may reference oracle/, the corpus encoder or the emulation."""
    pkg = os.path.join(ROOT, "wavpackdecoder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "librefdec" not in txt and "refdec.h" not in txt and "libwvb_emul" not in txt, f
