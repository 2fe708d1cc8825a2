"""Host-side access to ExtendedOctree files (the payload of a UVF TOC block) through the C ABI's host-only helpers
tvk_octree_file_probe / tvk_octree_file_read_brick -- no device needed.  Mirrors what a Tuvok client gets from
ExtendedOctree::Open / GetBrickData (IO/UVF/ExtendedOctree/ExtendedOctree.cpp:87-165,313-360)."""
import ctypes as C
import os

import numpy as np

from . import _lib as L

_NP = {L.U8: np.uint8, L.U16: np.uint16, L.F32: np.float32, L.RGBA8: np.dtype((np.uint8, 4))}


def probe(path, offset=0, uvf_file_version=5):
    """Parse header + table of contents.  Returns OctreeFileInfo; raises TvkError on a malformed file."""
    info = L.OctreeFileInfo()
    rc = L.lib().tvk_octree_file_probe(os.fsencode(path), int(offset), int(uvf_file_version), C.byref(info))
    if rc:
        raise L.TvkError(rc, (L.lib().tvk_last_error(None) or b"").decode())
    return info


def read_brick(path, x, y, z, lod, info=None, offset=0, uvf_file_version=5):
    """One decoded brick as ndarray [sz, sy, sx] (own size incl. ghost; [sz, sy, sx, 4] for a colour file)."""
    info = info or probe(path, offset, uvf_file_version)
    if info.dtype not in _NP:
        raise ValueError("component type not on the hot path")
    cap = int(np.prod(list(info.max_brick_size))) * np.dtype(_NP[info.dtype]).itemsize
    buf = np.zeros(cap, np.uint8)
    size = L.u32x3()
    rc = L.lib().tvk_octree_file_read_brick(os.fsencode(path), int(offset), int(uvf_file_version), x, y, z, lod,
                                            buf.ctypes.data_as(C.c_void_p), cap, size)
    if rc:
        raise L.TvkError(rc, (L.lib().tvk_last_error(None) or b"").decode())
    sx, sy, sz = (int(v) for v in size)
    out = np.frombuffer(buf.tobytes(), _NP[info.dtype], sx * sy * sz)
    return out.reshape((sz, sy, sx, 4) if info.dtype == L.RGBA8 else (sz, sy, sx))


def uvf_stats(path, timestep=0):
    """What UVFDataset derives from the MaxMin / histogram blocks: dict(range (lo, hi) or None, hist1d_size, hist1d_filled,
    max_gradient_magnitude, hist2d_size)."""
    rng = (C.c_double * 2)()
    n1, f1, mg, h2 = C.c_uint64(), C.c_uint64(), C.c_float(), (C.c_uint64 * 2)()
    lib = L.lib()
    rc = lib.tvk_uvf_probe_stats(os.fsencode(path), int(timestep), C.byref(rng), C.byref(n1), C.byref(f1), C.byref(mg), C.byref(h2))
    if rc:
        raise L.TvkError(rc, (lib.tvk_last_error(None) or b"").decode())
    return dict(range=(rng[0], rng[1]) if rng[1] >= rng[0] else None, hist1d_size=n1.value, hist1d_filled=f1.value,
                max_gradient_magnitude=mg.value, hist2d_size=(h2[0], h2[1]))


def uvf_probe(path, timestep=0):
    """Walk a .uvf container: dict(toc_payload_offset, file_version, n_blocks, n_timesteps, maxmin (n, 4) or None)."""
    off, ver, nb, nt, nm = (C.c_uint64() for _ in range(5))
    lib = L.lib()
    rc = lib.tvk_uvf_probe(os.fsencode(path), int(timestep), C.byref(off), C.byref(ver), C.byref(nb), C.byref(nt), None, 0,
                           C.byref(nm))
    if rc:
        raise L.TvkError(rc, (lib.tvk_last_error(None) or b"").decode())
    mm = None
    if nm.value:
        mm = np.zeros((nm.value, 4), np.float64)
        lib.tvk_uvf_probe(os.fsencode(path), int(timestep), None, None, None, None, mm.ctypes.data_as(C.c_void_p),
                          nm.value, None)
    return dict(toc_payload_offset=off.value, file_version=ver.value, n_blocks=nb.value, n_timesteps=nt.value, maxmin=mm)
