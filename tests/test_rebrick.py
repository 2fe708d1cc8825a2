"""DynamicBrickingDS (IO/DynamicBrickingDS.cpp, IOManager::LoadRebrickedDataset IO/IOManager.cpp:1281-1320):
a dataset converted with LARGE bricks is re-cut into the pool's small bricks when it is loaded.

  * the oracle's restatement (oracle/orc.py `Rebricked`) against the KATs of IO/test/rebricking.h,
  * the restatement against the UNMODIFIED reference class (oracle/_ref/ref_dynbrick: DynamicBrickingDS.cpp, BrickCache.cpp,
    BMinMax.cpp over the reference's own UVFDataset on a .uvf written by the reference's own UVF classes): LoD count, layouts,
    voxel counts, MM_PRECOMPUTE min / max and the voxels of every target brick, and the constructor's refusals,
  * (-m gpu) the device path `tvk_open_octree_file_rebricked` against the restatement on reference-written golden files:
    every brick and every min / max bit for bit, and frames identical to the same dataset served brick by brick
    from the oracle."""
import os
import subprocess

import numpy as np
import pytest

import tuvok_b200 as tb
from oracle import orc
from tuvok_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_UVF = os.path.join(ROOT, "oracle", "_ref", "ref_uvf")
REF_DYN = os.path.join(ROOT, "oracle", "_ref", "ref_dynbrick")


def _have_ref():
    if not (os.path.exists(REF_UVF) and os.path.exists(REF_DYN)) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(REF_UVF) and os.path.exists(REF_DYN)


needs_ref = pytest.mark.skipif(not _have_ref(), reason="oracle/_ref/ref_dynbrick not built (reference tree absent)")


def fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


# ---------------------------------------------------------------------------------------------- KATs of IO/test/rebricking.h
RAMP = np.arange(64, dtype=np.uint8).reshape(1, 8, 8)          # rebricking.h:12-21 (stored as 8 bit after range detection)


def ramp_octree():
    return orc.Octree(RAMP, 16, 2)                              # rebricking.h:29-35: brick 16, overlap 2


def test_kat_make_two_and_y():
    assert orc.Rebricked(ramp_octree(), (8, 16, 16)).total_bricks() == 5     # tmake_two, rebricking.h:119-124
    assert orc.Rebricked(ramp_octree(), (16, 8, 16)).total_bricks() == 5     # ty, :134-139


@pytest.mark.parametrize("bs", [(9, 16, 16), (9, 9, 16)])
def test_kat_uneven_throws(bs):                                 # tuneven / tuneven_multiple_dims, :127-131 / :141-145
    with pytest.raises(ValueError):
        orc.Rebricked(ramp_octree(), bs)


def test_kat_no_dynamic_and_data():
    r = orc.Rebricked(ramp_octree(), (16, 16, 16))              # tdata_simple / tdata_no_dynamic, :162-243
    assert r.brick_size(0, 0, 0, 0) == (12, 12, 5)
    d = r.brick(0, 0, 0, 0)
    assert np.array_equal(d[2, 2:10, 2:10], RAMP[0])
    assert np.array_equal(d, ramp_octree().brick(0, 0, 0, 0))   # tno_dynamic, :485-502: identical to the source


def test_kat_half_split():
    r = orc.Rebricked(ramp_octree(), (6, 16, 16))               # verify_half_split, :246-267; tvoxel_count, :277-296
    assert r.brick_size(0, 0, 0, 0) == (6, 12, 5) and r.brick_size(1, 0, 0, 0) == (6, 12, 5)
    d = r.brick(0, 0, 0, 0)
    assert np.array_equal(d[2, 2:10, 2:4], RAMP[0][:, 0:2])


def test_kat_minmax_includes_the_ghost():
    r = orc.Rebricked(ramp_octree(), (16, 8, 16))               # tminmax_dynamic, :392-416
    assert r.brick(0, 0, 0, 0).shape == (5, 8, 12)
    assert r.minmax(0, 0, 0, 0) == (0.0, 47.0)                  # "47: includes the ghost!"
    assert r.minmax(0, 1, 0, 0) == (0.0, 63.0)


def test_level0_equals_direct_bricking():
    # a target brick's ghost at a source-brick border is the source brick's ghost = the neighbouring voxels of the level, so the
    # finest level equals a direct conversion with the small brick size (no stale corners at level 0, orc_octree.c)
    vol = synth.synth_volume(synth.V_NOISE, (60, 52, 30), orc.U16, 0x5EED)
    r = orc.Rebricked(orc.Octree(vol, 28, 2), 12)
    direct = orc.Octree(vol, 12, 2)
    assert r.brick_count(0) == direct.brick_count(0)
    for (x, y, z, lod) in r.iter_bricks():
        if lod == 0:
            assert np.array_equal(r.brick(x, y, z, 0), direct.brick(x, y, z, 0))


# ---------------------------------------------------------------------------------------------- the unmodified reference class
def ref_dump(tmp_path, vol, dtype, brick, overlap, target, comp=1):
    raw = tmp_path / "in.raw"
    vol.tofile(raw)
    uvf = tmp_path / "vol.uvf"
    nz, ny, nx = vol.shape
    subprocess.check_call([REF_UVF, str(raw), str(uvf), {orc.U8: "u8", orc.U16: "u16"}[dtype], str(nx), str(ny), str(nz), str(brick),
                           str(overlap), str(comp), "0", "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = tmp_path / "dyn.txt"
    subprocess.check_call([REF_DYN, str(uvf), str(out)] + [str(t) for t in target] + ["256"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL, cwd=str(tmp_path))
    head, lods, bricks = {}, {}, {}
    for line in open(out):
        t = line.split()
        if t[0] == "throws":
            return {"throws": " ".join(t[1:])}, lods, bricks
        if t[0] == "lods":
            head.update(lods=int(t[1]), total=int(t[3]), maxbrick=tuple(int(v) for v in t[5:8]), maxused=tuple(int(v) for v in t[9:12]),
                        overlap=tuple(int(v) for v in t[13:16]), bits=int(t[17]))
        elif t[0] == "lod":
            lods[int(t[1])] = dict(domain=tuple(int(v) for v in t[3:6]), layout=tuple(int(v) for v in t[7:10]))
        elif t[0] == "brick":
            bricks[(int(t[1]), int(t[2]))] = dict(vox=tuple(int(v) for v in t[4:7]), mm=(float.fromhex(t[8]), float.fromhex(t[9])),
                                                  fnv=int(t[11], 16))
    return head, lods, bricks


REF_CASES = [
    (synth.V_NOISE, (8, 8, 1), orc.U8, 16, 2, (8, 16, 16)),           # the shape of rebricking.h
    (synth.V_NOISE, (44, 36, 28), orc.U8, 16, 2, (8, 8, 8)),          # ratio 3 on every axis
    (synth.V_SPH, (80, 70, 50), orc.U8, 36, 2, (20, 12, 36)),         # ratios 2 / 4 / 1, ragged last bricks
    (synth.V_NOISE, (60, 52, 30), orc.U16, 28, 2, (12, 12, 12)),      # 16 bit
    (synth.V_NOISE, (60, 52, 30), orc.U16, 28, 2, (64, 64, 64)),      # a target larger than the source: nothing is re-cut
]


@needs_ref
@pytest.mark.parametrize("kind,size,dtype,brick,overlap,target", REF_CASES)
def test_restatement_matches_reference_dynamic_bricking(tmp_path, kind, size, dtype, brick, overlap, target):
    vol = synth.synth_volume(kind, size, dtype, 0x5EED)
    # (IOManager.cpp:1301-1305 clamps the request before the class sees it; the class itself would assert)
    tgt = tuple(min(t, brick) for t in target)
    head, lods, bricks = ref_dump(tmp_path, vol, dtype, brick, overlap, tgt)
    src = orc.Octree(vol, brick, overlap)
    r = orc.Rebricked(src, target)
    assert head["lods"] == r.lod_count and head["total"] == r.total_bricks() and head["maxbrick"] == r.max_brick
    assert head["overlap"] == (overlap,) * 3
    inner = brick - 2 * overlap
    # levels whose SOURCE bricks the oracle does not restate (converter reads stale memory there, orc_octree.c "Q2")
    q2 = {lod for lod in range(src.lod_count)
          if any(0 < (src.lod_size(lod)[a] % inner) < overlap and src.brick_count(lod)[a] > 1 for a in range(3))}
    n = 0
    for (x, y, z, lod) in r.iter_bricks():
        bc = r.brick_count(lod)
        assert lods[lod]["layout"] == bc and lods[lod]["domain"] == r.lod_size(lod)
        b = bricks[(lod, x + bc[0] * (y + bc[1] * z))]
        assert b["vox"] == r.brick_size(x, y, z, lod), (x, y, z, lod)
        if lod in q2:
            continue
        assert b["mm"] == r.minmax(x, y, z, lod), (x, y, z, lod)
        assert b["fnv"] == fnv1a(np.ascontiguousarray(r.brick(x, y, z, lod)).tobytes()), (x, y, z, lod)
        n += 1
    assert n > 0 and len(bricks) == r.total_bricks()
    assert head["maxused"] == tuple(max(b["vox"][a] for b in bricks.values()) for a in range(3))


@needs_ref
def test_reference_refuses_what_the_restatement_refuses(tmp_path):
    vol = synth.synth_volume(synth.V_NOISE, (44, 36, 28), orc.U8, 0x5EED)
    head, _, _ = ref_dump(tmp_path, vol, orc.U8, 16, 2, (9, 16, 16))
    assert "integer multiple" in head["throws"]
    with pytest.raises(ValueError, match="integer multiple"):
        orc.Rebricked(orc.Octree(vol, 16, 2), (9, 16, 16))


# ---------------------------------------------------------------------------------------------- the device path
GPU_CASES = [
    ("octree_u8_b36_zlib", (20, 12, 36)),
    ("octree_u8_b36_zlib", (12, 12, 12)),
    ("octree_u16_b28_lz4", (12, 12, 12)),
    ("octree_u16_b28_lz4", (16, 28, 10)),
    ("octree_u16_lz4_morton", (8, 8, 8)),
    ("octree_f32_none", (8, 8, 8)),
    ("octree_u16_b28_lz4", (64, 64, 64)),          # clamped to the source's brick size: a plain load
]


def _golden_case(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_octree_golden", os.path.join(GOLDEN, "make_octree_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.CASES[name], m.volume(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name,target", GPU_CASES)
def test_device_rebricking_matches_restatement(name, target):
    (kind, size, dt, _, brick, ov, _, _), vol = _golden_case(name)
    src = orc.Octree(vol, brick, ov)
    want = orc.Rebricked(src, target)
    r = tb.CudaGridLeaper(max_gpu_mem=64 << 20)
    info = r.OpenRebrickedOctreeFile(os.path.join(GOLDEN, name + ".bin"), target)
    assert info.brick_count == src.total_bricks                      # the FILE's header
    gi = r.info()
    assert gi.lod_count == want.lod_count
    mm = r.minmax(want.total_bricks())
    inner = brick - 2 * ov
    q2 = {lod for lod in range(src.lod_count)
          if any(0 < (src.lod_size(lod)[a] % inner) < ov and src.brick_count(lod)[a] > 1 for a in range(3))}
    i = 0
    for (x, y, z, lod) in want.iter_bricks():
        assert r.brick_size(x, y, z, lod) == want.brick_size(x, y, z, lod)
        if lod not in q2:
            b = want.brick(x, y, z, lod)
            assert np.array_equal(r.brick(x, y, z, lod, dt), b), (x, y, z, lod)
            assert (mm[i, 0], mm[i, 1]) == (float(b.min()), float(b.max())), (x, y, z, lod)
        i += 1
    assert i == len(mm)
    r.Cleanup()


@pytest.mark.gpu
def test_rebricked_file_renders_like_the_same_bricks_served_by_callback():
    from scene import Scene
    name, target = "octree_u8_b36_zlib", (20, 20, 20)
    (kind, size, dt, _, brick, ov, _, _), vol = _golden_case(name)
    want = orc.Rebricked(orc.Octree(vol, brick, ov), target)
    s = Scene(kind=kind, size=size, dtype=dt, brick=target[0], overlap=ov, width=72, height=56, lighting=True,
              rotation=(tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32), tf_center=0.25, tf_inv_gradient=0.3,
              seed=0x5EED)
    a = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    a.OpenRebrickedOctreeFile(os.path.join(GOLDEN, name + ".bin"), target, range_max=s.range_max, max_gradient_magnitude=s.max_grad)
    n = want.total_bricks()
    table = a.minmax(n)
    b = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    b.RegisterDataset(size, target, ov, dt, table, lambda x, y, z, lod: want.brick(x, y, z, lod), range_max=s.range_max,
                      max_gradient_magnitude=s.max_grad)
    imgs = []
    for r in (a, b):
        r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
        r.Resize(s.width, s.height); r.SetRotation(s.rotation)
        r.CreateVolumePool(s._pool_size)
        assert r.PaintUntilConverged().converged
        imgs.append(r.ReadRGBA32F())
    assert imgs[0][..., 3].max() > 0.1
    assert np.array_equal(imgs[0], imgs[1])
    assert np.array_equal(a.page_table(), b.page_table())
    a.Cleanup(); b.Cleanup()


@pytest.mark.gpu
def test_device_rebricking_refuses_a_non_divisor():
    r = tb.CudaGridLeaper(max_gpu_mem=64 << 20)
    with pytest.raises(tb.TvkError, match="multiple"):
        r.OpenRebrickedOctreeFile(os.path.join(GOLDEN, "octree_u8_b36_zlib.bin"), (9, 36, 36))
    r.Cleanup()
