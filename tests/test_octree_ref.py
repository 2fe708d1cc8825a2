"""The oracle's bricker / LOD pyramid / min-max against the UNMODIFIED reference ExtendedOctreeConverter
(oracle/_ref/ref_octree, compiled from /root/reference by oracle/Makefile).  Bit-exact, every brick of
every LOD, every voxel incl. ghost cells, and the per-brick min/max."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import orc

DT = {orc.U8: ("u8", np.uint8), orc.U16: ("u16", np.uint16), orc.F32: ("f32", np.float32)}


def run_reference(binary, tmp_path, vol, dtype, brick, overlap, clamp, median=False):
    name, npdt = DT[dtype]
    raw = tmp_path / "in.raw"
    out = tmp_path / "out.bin"
    vol.astype(npdt).tofile(raw)
    nz, ny, nx = vol.shape
    subprocess.check_call([binary, str(raw), str(out), name, str(nx), str(ny), str(nz), str(brick), str(overlap),
                           str(int(clamp)), str(int(median))], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    data = out.read_bytes()
    lods, total = struct.unpack_from("<QQ", data, 0)
    off = 16
    bricks = []
    es = np.dtype(npdt).itemsize
    for _ in range(total):
        sx, sy, sz = struct.unpack_from("<QQQ", data, off); off += 24
        mn, mx = struct.unpack_from("<dd", data, off); off += 16
        n = sx * sy * sz
        vox = np.frombuffer(data, npdt, n, off).reshape(sz, sy, sx); off += n * es
        bricks.append(((sx, sy, sz), mn, mx, vox))
    return lods, bricks


def rand_volume(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    if dtype == orc.F32:
        return rng.random(shape, dtype=np.float32)
    hi = 255 if dtype == orc.U8 else 65535
    return rng.integers(1, hi, size=shape, endpoint=True).astype(DT[dtype][1])   # >= 1: zeros only from borders


CASES = [
    # (nz, ny, nx), dtype, brick, overlap, clamp
    ((8, 8, 8), orc.U8, 8, 2, False),
    ((16, 16, 16), orc.U8, 8, 2, False),
    ((24, 20, 28), orc.U16, 12, 2, False),
    ((33, 17, 40), orc.U16, 12, 2, False),
    ((32, 32, 32), orc.F32, 12, 2, False),
    ((1, 8, 8), orc.U8, 16, 2, False),          # the rebricking.h shape
    ((40, 40, 40), orc.U8, 20, 2, False),
    ((36, 36, 36), orc.U16, 10, 1, False),
    ((30, 26, 22), orc.U8, 14, 3, False),
]


@pytest.mark.parametrize("shape,dtype,brick,overlap,clamp", CASES)
def test_oracle_matches_reference_converter(ref_octree_bin, tmp_path, shape, dtype, brick, overlap, clamp):
    vol = rand_volume(shape, dtype, seed=sum(shape) + brick)
    lods, ref = run_reference(ref_octree_bin, tmp_path, vol, dtype, brick, overlap, clamp)
    o = orc.Octree(vol, brick, overlap, clamp=clamp)
    assert o.lod_count == lods
    assert o.total_bricks == len(ref)
    inner = brick - 2 * overlap
    for (x, y, z, lod) in o.iter_bricks():
        i = o.brick_index(x, y, z, lod)
        size, mn, mx, vox = ref[i]
        assert o.brick_size(x, y, z, lod) == size, (x, y, z, lod)
        # outside the contract (orc_octree.c header, "Q2"): a last brick whose remainder is smaller than the
        # overlap makes the reference read stale memory
        ls = o.lod_size(lod)
        q2 = any(0 < (ls[a] % inner) < overlap and o.brick_count(lod)[a] > 1 for a in range(3))
        if q2:
            continue
        mine = o.brick(x, y, z, lod)
        assert np.array_equal(mine, vox), "brick %s differs" % ((x, y, z, lod),)
        assert (o.minmax[i, 0], o.minmax[i, 1]) == (mn, mx), (x, y, z, lod)


@pytest.mark.parametrize("shape,dtype,brick,overlap", [((24, 20, 28), orc.U16, 12, 2), ((33, 17, 40), orc.U8, 12, 2),
                                                       ((32, 32, 32), orc.F32, 12, 2), ((1, 8, 8), orc.U8, 16, 2)])
def test_oracle_median_pyramid_matches_reference_converter(ref_octree_bin, tmp_path, shape, dtype, brick, overlap):
    """bComputeMedian = true (ExtendedOctreeConverter.inc:1-248, VolumeTools.h:168-262): the coarser levels hold medians"""
    vol = rand_volume(shape, dtype, seed=sum(shape) + 3 * brick)
    lods, ref = run_reference(ref_octree_bin, tmp_path, vol, dtype, brick, overlap, False, median=True)
    o = orc.Octree(vol, brick, overlap, median=True)
    mean = orc.Octree(vol, brick, overlap)
    assert o.lod_count == lods and o.total_bricks == len(ref)
    inner = brick - 2 * overlap
    differs = False
    for (x, y, z, lod) in o.iter_bricks():
        ls = o.lod_size(lod)
        if any(0 < (ls[a] % inner) < overlap and o.brick_count(lod)[a] > 1 for a in range(3)):
            continue
        i = o.brick_index(x, y, z, lod)
        assert np.array_equal(o.brick(x, y, z, lod), ref[i][3]), (x, y, z, lod)
        assert (o.minmax[i, 0], o.minmax[i, 1]) == (ref[i][1], ref[i][2])
        differs = differs or (lod > 0 and not np.array_equal(o.brick(x, y, z, lod), mean.brick(x, y, z, lod)))
    assert differs or lods == 1            # the median pyramid is not the mean pyramid


def test_ghost_corner_quirk_is_reference_behaviour(ref_octree_bin, tmp_path):
    """FillOverlap's copy order leaves three ghost corners of interior LOD>=1 bricks zero (Q1 in
    orc_octree.c).  Show it on the reference itself so the restatement is not an invention."""
    vol = rand_volume((48, 48, 48), orc.U8, 7)
    _, ref = run_reference(ref_octree_bin, tmp_path, vol, orc.U8, 12, 2, False)
    o = orc.Octree(vol, 12, 2)
    bc = o.brick_count(1)
    assert min(bc) >= 3
    i = o.brick_index(1, 1, 1, 1)                      # interior brick of LOD 1
    vox = ref[i][3]
    assert not vox[0:2, 10:12, 10:12].any()            # (right, bottom, front)
    assert not vox[10:12, 0:2, 10:12].any()            # (right, top, back)
    assert not vox[10:12, 10:12, 0:2].any()            # (left, bottom, back)
    assert vox[10:12, 10:12, 10:12].all()              # the opposite corner is filled


@pytest.mark.parametrize("shape,brick,overlap", [((24, 20, 28), 12, 2), ((33, 17, 40), 12, 2), ((40, 36, 44), 16, 2)])
def test_colour_octree_matches_reference_multi_component_converter(ref_octree_bin, tmp_path, shape, brick, overlap):
    """Colour (RGBA8) volumes: orc.ColorOctree -- four scalar conversions interleaved, min / max of the alpha component --
    against the UNMODIFIED converter run with iComponentCount = 4 on the interleaved volume: every brick, every voxel of every
    component incl. ghost cells, and the alpha statistics a renderer sees (uvfDataset.cpp:1188)."""
    rng = np.random.default_rng(sum(shape) + brick)
    vol = rng.integers(1, 255, size=shape + (4,), endpoint=True).astype(np.uint8)
    raw, out = tmp_path / "in.raw", tmp_path / "out.bin"
    vol.tofile(raw)
    nz, ny, nx = shape
    subprocess.check_call([ref_octree_bin, str(raw), str(out), "rgba8", str(nx), str(ny), str(nz), str(brick), str(overlap), "0", "0"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    data = out.read_bytes()
    lods, total = struct.unpack_from("<QQ", data, 0)
    off = 16
    ref = []
    for _ in range(total):
        sx, sy, sz = struct.unpack_from("<QQQ", data, off); off += 24
        mn, mx = struct.unpack_from("<dd", data, off); off += 16
        n = sx * sy * sz * 4
        ref.append(((sx, sy, sz), mn, mx, np.frombuffer(data, np.uint8, n, off).reshape(sz, sy, sx, 4))); off += n
    o = orc.ColorOctree(vol, brick, overlap)
    assert o.lod_count == lods and o.total_bricks == len(ref)
    inner = brick - 2 * overlap
    checked = 0
    for (x, y, z, lod) in o.iter_bricks():
        i = o.brick_index(x, y, z, lod)
        size, mn, mx, vox = ref[i]
        assert o.brick_size(x, y, z, lod) == size
        ls = o.lod_size(lod)
        if any(0 < (ls[a] % inner) < overlap and o.brick_count(lod)[a] > 1 for a in range(3)):
            continue                                    # "Q2", as for scalar volumes
        assert np.array_equal(o.brick(x, y, z, lod), vox), (x, y, z, lod)
        assert (o.minmax[i, 0], o.minmax[i, 1]) == (mn, mx), (x, y, z, lod)
        checked += 1
    assert checked >= 3
