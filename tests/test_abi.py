"""The C-ABI library loads, exports every symbol include/tvk.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import tuvok_b200 as tb
from tuvok_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tvk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvk_[a-z0-9_]+)\s*\(", src)) - {"tvk_brick_cb", "tvk_log_cb"})


def test_library_is_built_in_tree():
    assert os.path.exists(tb.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.dirname(tb.LIB_PATH).endswith("tuvok_b200")


def test_every_declared_symbol_is_exported_and_bound():
    lib = tb.lib()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), "libtvkcuda.so does not export %s" % s
        assert s in L.SIGNATURES, "tuvok_b200/_lib.py does not bind %s" % s
    assert sorted(L.SIGNATURES) == syms
    assert lib.tvk_abi_version() == 1


def test_struct_layouts_match_the_header():
    # sizes implied by include/tvk.h on LP64
    assert C.sizeof(L.DeviceCfg) == 32
    assert C.sizeof(L.RenderParams) == 272
    assert C.sizeof(L.VolumeDesc) == 80
    assert C.sizeof(L.Info) == 8 + 8 + 16 * 12 * 2 + 64 + 36 + 4 + 8
    assert C.sizeof(L.FrameStats) == 4 * 3 + 4 + 8 * 7 + 16


def test_sass_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % tb.LIB_PATH).read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback_without_device():
    lib = tb.lib()
    h = C.c_void_p()
    rc = lib.tvk_create(None, C.byref(h))
    assert rc == L.ERR_NO_DEVICE and not h
    assert b"no CPU fallback" in lib.tvk_last_error(None)
    with pytest.raises(tb.TvkError):
        tb.CudaGridLeaper()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tuvok_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "orc.h" not in txt, f
                assert "liborc" not in txt, f
