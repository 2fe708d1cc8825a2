"""Procedural multi-resolution dataset (BASELINE configs[4]), host side: the generator threads of the library
(tvk_procedural_brick = what they write into the pinned staging memory) against the numpy statement of the same field
(tuvok_b200/synth.py), brick by brick and bit for bit; the geometry against BASELINE.md's numbers for 8192^3 / 128^3."""
import ctypes as C

import numpy as np
import pytest

from tuvok_b200 import _lib as L
from tuvok_b200 import synth

NP = {L.U8: np.uint8, L.U16: np.uint16, L.F32: np.float32}


def lib_brick(kind, size, dtype, brick, overlap, x, y, z, lod, seed=0x5EED):
    buf = np.zeros(brick[0] * brick[1] * brick[2], NP[dtype])
    out = L.u32x3()
    rc = L.lib().tvk_procedural_brick(kind, L.u32x3(*size), dtype, seed, L.u32x3(*brick), overlap, x, y, z, lod,
                                      buf.ctypes.data_as(C.c_void_p), buf.nbytes, out)
    assert rc == L.OK
    bs = tuple(out)
    return buf[:bs[0] * bs[1] * bs[2]].reshape(bs[2], bs[1], bs[0])


def test_c5_geometry_is_baselines():
    """8192^3 u8 in 128^3 bricks -> 345 870 bricks in 8 LoDs"""
    n, lods = C.c_uint64(), C.c_uint32()
    assert L.lib().tvk_procedural_brick_count(L.u32x3(8192, 8192, 8192), L.u32x3(128, 128, 128), 2, C.byref(n), C.byref(lods)) == L.OK
    assert (n.value, lods.value) == (345870, 8)
    sizes, layouts, offs = synth.procedural_geometry((8192,) * 3, 128, 2)
    assert offs[-1] == 345870 and len(sizes) == 8 and layouts[0] == (67, 67, 67) and sizes[-1] == (64, 64, 64)


@pytest.mark.parametrize("kind", [synth.V_SPH, synth.V_NOISE, synth.V_RAMP])
@pytest.mark.parametrize("dtype", [L.U8, L.U16, L.F32])
def test_generator_equals_numpy_field(kind, dtype):
    size, brick, ov = (100, 70, 90), (36, 28, 20), 2          # ragged last bricks on every axis, anisotropic bricks
    sizes, layouts, offs = synth.procedural_geometry(size, brick, ov)
    n, lods = C.c_uint64(), C.c_uint32()
    assert L.lib().tvk_procedural_brick_count(L.u32x3(*size), L.u32x3(*brick), ov, C.byref(n), C.byref(lods)) == L.OK
    assert n.value == offs[-1] and lods.value == len(sizes)
    for lod, lay in enumerate(layouts):
        for co in {(0, 0, 0), (lay[0] - 1, lay[1] - 1, lay[2] - 1), (lay[0] // 2, lay[1] // 2, lay[2] // 2), (lay[0] - 1, 0, lay[2] // 2)}:
            ref = synth.procedural_brick(kind, size, dtype, 0x5EED, brick, ov, *co, lod)
            got = lib_brick(kind, size, dtype, brick, ov, *co, lod)
            assert got.shape == ref.shape and np.array_equal(got, ref), (lod, co)


def test_level_zero_is_the_bricked_synthetic_volume():
    """level 0 of the procedural hierarchy is exactly tvk_synth_volume's field, so a level-0 brick is the corresponding
    window of the full volume (ghost voxels = the neighbours' voxels, 0 outside)"""
    size = (90, 64, 50)
    vol = synth.synth_volume(synth.V_NOISE, size, L.U16)
    b = lib_brick(synth.V_NOISE, size, L.U16, (36, 36, 36), 2, 1, 1, 0, 0)
    assert np.array_equal(b[2:, :34, :], vol[0:34, 30:64, 30:66]) and not b[:2].any() and not b[:, 34:, :].any()


def test_big_brick_of_the_512_gib_volume_and_seed():
    size, brick = (8192,) * 3, (128,) * 3
    a = lib_brick(synth.V_NOISE, size, L.U8, brick, 2, 30, 33, 35, 0)
    ref = synth.procedural_brick(synth.V_NOISE, size, L.U8, 0x5EED, 128, 2, 30, 33, 35, 0)
    assert np.array_equal(a, ref) and a.any()
    b = lib_brick(synth.V_NOISE, size, L.U8, brick, 2, 30, 33, 35, 0, seed=7)
    assert not np.array_equal(a, b)
    # outside the field's support (corner of the domain) a brick is empty
    assert not lib_brick(synth.V_NOISE, size, L.U8, brick, 2, 0, 0, 0, 0).any()


def test_bad_arguments_are_refused():
    n = C.c_uint64()
    assert L.lib().tvk_procedural_brick_count(L.u32x3(64, 64, 0), L.u32x3(36, 36, 36), 2, C.byref(n), None) != L.OK
    assert L.lib().tvk_procedural_brick_count(L.u32x3(64, 64, 64), L.u32x3(4, 36, 36), 2, C.byref(n), None) != L.OK
    buf = np.zeros(8, np.uint8)
    assert L.lib().tvk_procedural_brick(1, L.u32x3(64, 64, 64), L.U8, 1, L.u32x3(36, 36, 36), 2, 0, 0, 0, 0,
                                        buf.ctypes.data_as(C.c_void_p), buf.nbytes, L.u32x3()) != L.OK      # buffer too small
    big = np.zeros(36 ** 3, np.uint8)
    assert L.lib().tvk_procedural_brick(1, L.u32x3(64, 64, 64), L.U8, 1, L.u32x3(36, 36, 36), 2, 2, 0, 0, 0,
                                        big.ctypes.data_as(C.c_void_p), big.nbytes, L.u32x3()) != L.OK      # x out of range
