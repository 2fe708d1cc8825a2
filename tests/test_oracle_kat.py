"""The oracle against the reference's OWN known-answer vectors (IO/test/rebricking.h) and geometry facts.
These pin the data side (bricking, ghost cells, min/max-with-ghost) of the oracle."""
import numpy as np
import pytest

from oracle import orc

# IO/test/rebricking.h:12-21 -- the 8x8x1 ramp 0..63 (the reference stores it as 8-bit after range detection)
RAMP = np.arange(64, dtype=np.uint8).reshape(1, 8, 8)


def test_single_brick_size_and_content():
    # rebricking.h:29-35 converts with brick 16 / overlap 2; :223-225 asserts 12x12x5
    o = orc.Octree(RAMP, 16, 2)
    assert o.brick_count(0) == (1, 1, 1)
    assert o.brick_size(0, 0, 0, 0) == (12, 12, 5)
    b = o.brick(0, 0, 0, 0)
    assert b.shape == (5, 12, 12)
    # rebricking.h:176-187: interior of slice z=2 equals the source; ghost = 2 on every side
    assert np.array_equal(b[2, 2:10, 2:10], RAMP[0])
    ghost = b.copy()
    ghost[2, 2:10, 2:10] = 0
    assert not ghost.any()          # zero border (bClampToEdge = false)


def test_half_split_bricks():
    # rebricking.h:246-267: bricks of 6x12x5 when split in x (max brick 6 = 2 inner + 4 ghost)
    o = orc.Octree(RAMP, (6, 16, 16), 2)
    assert o.brick_count(0) == (4, 1, 1)
    assert o.brick_size(0, 0, 0, 0) == (6, 12, 5)
    b = o.brick(0, 0, 0, 0)
    # interior columns 0..1 of the source + 2 ghost columns from the right neighbour
    assert np.array_equal(b[2, 2:10, 2:6], RAMP[0][:, 0:4])
    assert not b[2, 2:10, 0:2].any()


def test_minmax_includes_ghost():
    # rebricking.h:392-416: split in Y at 8 => brick0 [0,47], brick1 [0,63] ("47: includes the ghost!")
    o = orc.Octree(RAMP, (16, 8, 16), 2)
    assert o.brick_count(0) == (1, 2, 1)
    assert o.brick_size(0, 0, 0, 0) == (12, 8, 5)
    mm = o.minmax
    i0, i1 = o.brick_index(0, 0, 0, 0), o.brick_index(0, 1, 0, 0)
    assert (mm[i0, 0], mm[i0, 1]) == (0.0, 47.0)
    assert (mm[i1, 0], mm[i1, 1]) == (0.0, 63.0)
    # gradient range never culls for TOC datasets (MaxMinDataBlock.cpp:175-195)
    assert mm[i0, 2] == -np.finfo(np.float64).max and mm[i0, 3] == np.finfo(np.float64).max


def test_range_detection_values():
    # rebricking.h:149-159: range 0..63
    o = orc.Octree(RAMP, 16, 2)
    assert o.minmax[:, 1].max() == 63.0


@pytest.mark.parametrize("vol,brick,ov,lods,bricks0", [
    ((512, 512, 512), 36, 2, 10, (16, 16, 16)),      # BASELINE C2: 4681 bricks in 5 pool LoDs
    ((256, 256, 256), 260, 2, 9, (1, 1, 1)),         # BASELINE C1: single brick
    ((100, 60, 33), 20, 2, 8, (7, 4, 3)),
])
def test_geometry(vol, brick, ov, lods, bricks0):
    # ExtendedOctree::ComputeMetadata: LOD sizes ceil(prev/2) until 1^3; bricks = ceil(size / inner)
    z = np.zeros((4, 4, 4), np.uint8)
    L = orc.lib()
    h = L.orc_octree_new(orc.u32x3(*vol), orc.u32x3(brick, brick, brick), ov, orc.U8)
    try:
        assert L.orc_octree_lod_count(h) == lods
        o = orc.u32x3()
        L.orc_octree_brick_count(h, 0, o)
        assert tuple(o) == bricks0
        del z
    finally:
        L.orc_octree_free(h)


def test_c2_c3_brick_totals():
    # C2 4 681 bricks / 5 pool LoDs; C3 299 593 bricks / 7 pool LoDs, page table 1.17 MB
    L = orc.lib()
    for n, total, single in ((512, 4681, 4), (2048, 299593, 6)):
        h = L.orc_octree_new(orc.u32x3(n, n, n), orc.u32x3(36, 36, 36), 2, orc.U16)
        single_lod = L.orc_octree_largest_single_brick_lod(h)
        assert single_lod == single
        idx_last = L.orc_octree_brick_index(h, 0, 0, 0, single_lod)
        assert idx_last + 1 == total
        L.orc_octree_free(h)
