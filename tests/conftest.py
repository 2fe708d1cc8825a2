import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checker():
    """The CPU oracle (test infrastructure) is compiled on demand; the product library must already be
    built by __graft_entry__.build() (it is NOT built here: tests must not hide a missing product)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "liborc.so"],
                          stdout=subprocess.DEVNULL)
    yield


@pytest.fixture(scope="session")
def ref_octree_bin():
    """oracle/_ref/ref_octree = the reference's own ExtendedOctreeConverter compiled from /root/reference."""
    path = os.path.join(ROOT, "oracle", "_ref", "ref_octree")
    if not os.path.exists(path) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/ref_octree not built (reference tree absent)")
    return path
