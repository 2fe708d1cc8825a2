"""The reference-side half of the boundary (integration/CUDAGridLeaper.{h,cpp}, the AbstrRenderer subclass INTEGRATION.md
describes) is COMPILED against the reference's own headers where the reference tree is mounted: every pure virtual of
Renderer/AbstrRenderer.h:112-881 is overridden (the file instantiates the class), every AbstrRenderer member it reads
exists with that type, and every tvk_* call matches include/tvk.h.  Tuvok cannot be linked in this image (Qt, GL and
bison are absent), so this is a compile of the one translation unit -- no link, nothing executed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Renderer")), reason="reference tree not mounted")

FLAGS = ["-std=c++11", "-w", "-DTUVOK_NO_QT=1", "-DGLEW_NO_GLU", "-I" + REF, "-I" + REF + "/Basics", "-I" + REF + "/IO",
         "-I" + REF + "/IO/3rdParty/boost", "-I" + REF + "/IO/3rdParty", "-I" + REF + "/IO/exception",
         "-I" + REF + "/3rdParty/GLEW", "-I" + REF + "/3rdParty/LUA", "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "integration")]


def test_shim_compiles_against_the_reference_headers(tmp_path):
    obj = os.path.join(str(tmp_path), "CUDAGridLeaper.o")
    p = subprocess.run(["g++", "-c"] + FLAGS + [os.path.join(ROOT, "integration", "CUDAGridLeaper.cpp"), "-o", obj],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-4000:]
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    # the factory that instantiates the class (so it is not abstract) and the calls into the C ABI it leaves undefined
    assert "tuvok::NewCUDAGridLeaper(" in syms
    for f in ("tvk_create", "tvk_set_volume", "tvk_create_pool", "tvk_render", "tvk_set_clip_plane", "tvk_pick",
              "tvk_sortlast_frame", "tvk_set_tf1d", "tvk_set_tf2d", "tvk_read_rgba8"):
        assert " U " + f in syms, f


def test_an_incomplete_shim_is_rejected(tmp_path):
    """the check has teeth: a subclass that leaves AbstrRenderer's pure virtuals open does not compile"""
    src = os.path.join(str(tmp_path), "bad.cpp")
    with open(src, "w") as f:
        f.write('#include "Renderer/AbstrRenderer.h"\nnamespace tuvok { struct Bad : AbstrRenderer { Bad() : AbstrRenderer(0, false, false, false) {} };\n'
                'AbstrRenderer* make() { return new Bad(); } }\n')
    p = subprocess.run(["g++", "-fsyntax-only"] + FLAGS + [src], capture_output=True, text=True)
    assert p.returncode != 0 and "abstract" in p.stderr
