"""Page table / visibility / LRU paging of the oracle restatement (oracle/orc_pool.cpp) against the UNMODIFIED
reference GLVolumePool.cpp (oracle/_ref/ref_pool: compiled in place from /root/reference over a recording
null-GL, see oracle/ref_shim/).  Everything is integer work => bit-exact:
  * the R32UI metadata texture as the reference uploaded it (what its shader reads) == oracle page table,
  * visibility counts, slot table (brick id, creation time, position) in the reference's own sorted order,
  * the brick-pool atlas as filled by the reference's glTexSubImage3D calls == atlas replayed from the
    oracle's slot decisions.
This pins SURVEY 8a8 / a9 / a10 to reference code instead of to a restatement."""
import numpy as np
import pytest

import pool_ref
from oracle import orc
from tuvok_b200 import synth

pytestmark = pytest.mark.skipif(not pool_ref.have_ref_pool(), reason="oracle/_ref/ref_pool not built (reference tree absent)")

VIS = {orc.RM_1DTRANS: "vis1d", orc.RM_2DTRANS: "vis2d", orc.RM_ISOSURFACE: "visiso"}


def vis_op(mode, args):
    n = {orc.RM_1DTRANS: 2, orc.RM_2DTRANS: 4, orc.RM_ISOSURFACE: 1}[mode]
    return (VIS[mode],) + tuple(args[:n])


def replay(tmp_path, vol, dtype, brick, overlap, pool_size, script, max3d=16384):
    """script: list of ("first",) | ("vis", mode, args) | ("upload", ids).  Runs it through the oracle and the
    reference (a dump after every step) and compares every observable."""
    size = (vol.shape[2], vol.shape[1], vol.shape[0])
    o = orc.Octree(vol, brick, overlap)
    lods = o.largest_single_brick_lod + 1
    pool = orc.Pool(pool_size, size, (brick,) * 3, overlap, lods, o.minmax, max_3d_dim=max3d)
    ops = []
    for st in script:
        if st[0] == "vis":
            ops.append(vis_op(st[1], st[2]))
        elif st[0] == "upload":
            ops.append(("upload", st[1]))
        else:
            ops.append(("first",))
        ops.append(("dump",))
    ref = pool_ref.run(tmp_path, o, size, brick, overlap, dtype, pool_size, ops, max3d=max3d)

    c = ref.create
    assert c["total"] == pool.total_bricks and c["lods"] == lods
    assert c["capacity"] == pool.capacity and c["metadim"] == pool.meta_dim
    assert c["offsets"] == [int(v) for v in pool.lod_offsets]

    ps = pool.pool_size
    atlas = np.zeros((ps[2], ps[1], ps[0]), orc.NP_DTYPE[dtype])

    def put(coord, key):
        cap = pool.capacity
        sx, sy, sz = coord % cap[0], (coord // cap[0]) % cap[1], coord // (cap[0] * cap[1])
        b = o.brick(*[int(v) for v in key])
        atlas[sz * brick:sz * brick + b.shape[0], sy * brick:sy * brick + b.shape[1], sx * brick:sx * brick + b.shape[2]] = b

    ev = iter(ref.events)
    for step, st in enumerate(script):
        if st[0] == "vis":
            counts = pool.recompute_visibility(st[1], *(tuple(st[2]) + (0.0,) * 4)[:4])
            e = next(ev)
            assert e == ("counts", counts), "step %d" % step
        elif st[0] == "upload":
            n, slots = pool.upload_bricks(np.array(st[1], np.uint32))
            e = next(ev)
            assert e == ("paged", n), "step %d" % step
            for key, sl in zip(st[1], slots):
                if sl != 0xFFFFFFFF:
                    put(int(sl), key)
        else:
            sl = pool.upload_first()
            assert next(ev) == ("first",)
            cap = pool.capacity
            put(cap[0] * cap[1] * cap[2] - 1, (0, 0, 0, lods - 1))
        kind, d = next(ev)
        assert kind == "dump"
        assert d["texture_equals_cpu"], "step %d: the reference's own texture and CPU table disagree" % step
        assert np.array_equal(d["meta"], pool.meta), "step %d: page table" % step
        ids, times, pos = pool.slots()
        assert np.array_equal(d["slot_brick"], ids), "step %d: slot bricks" % step
        assert np.array_equal(d["slot_time"], times), "step %d: slot times" % step
        assert np.array_equal(d["slot_pos"], pos), "step %d: slot positions" % step
    last = ref.dumps()[-1]
    assert np.array_equal(last["atlas"], atlas), "pool atlas contents"
    return pool


def all_keys(vol, brick, overlap):
    o = orc.Octree(vol, brick, overlap)
    return [k for k in o.iter_bricks(max_lod=o.largest_single_brick_lod)]


CASES = [
    # kind, (x, y, z), dtype, brick, overlap, pool (texels), seed
    (synth.V_SPH, (64, 64, 64), orc.U8, 20, 2, (60, 40, 40), 1),        # 12 slots: constant eviction
    (synth.V_NOISE, (100, 90, 70), orc.U16, 20, 2, (100, 80, 60), 2),   # odd layouts in every axis, 60 slots
    (synth.V_SPH, (48, 48, 48), orc.F32, 12, 2, (48, 48, 36), 3),
    (synth.V_NOISE, (75, 41, 130), orc.U8, 14, 3, (56, 56, 42), 4),     # 3-voxel ghost, ragged
    (synth.V_SPH, (36, 36, 36), orc.U16, 10, 1, (30, 30, 30), 5),       # 1-voxel ghost
]


@pytest.mark.parametrize("kind,size,dtype,brick,overlap,pool_size,seed", CASES)
@pytest.mark.parametrize("mode", [orc.RM_1DTRANS, orc.RM_2DTRANS, orc.RM_ISOSURFACE])
def test_restatement_matches_reference_pool(tmp_path, kind, size, dtype, brick, overlap, pool_size, seed, mode):
    vol = synth.synth_volume(kind, size, dtype, 0x5EED + seed)
    rng = np.random.default_rng(seed * 31 + mode)
    keys = all_keys(vol, brick, overlap)
    top = {orc.U8: 255.0, orc.U16: 65535.0, orc.F32: 1.0}[dtype]

    def vis_args():
        if mode == orc.RM_ISOSURFACE:
            return (float(rng.uniform(0.05, 0.95) * top),)
        lo = float(rng.uniform(0.0, 0.7) * top)
        hi = float(min(top, lo + rng.uniform(0.01, 0.5) * top))
        if mode == orc.RM_2DTRANS:
            return (lo, hi, 0.0, 255.0)
        return (lo, hi)

    script = [("first",), ("vis", mode, vis_args())]
    for rnd in range(10):
        pick = rng.choice(len(keys), size=int(rng.integers(1, 14)), replace=False)
        script.append(("upload", [keys[i] for i in pick]))
        if rnd in (3, 6):
            script.append(("vis", mode, vis_args()))
    replay(tmp_path, vol, dtype, brick, overlap, pool_size, script)


def test_visibility_before_first_brick_and_empty_tf(tmp_path):
    """TF with no opaque entry (ComputeNonZeroLimits -> lo = n, hi = 0): everything is empty; then a fully
    opaque TF: nothing is; then mode switches on the same pool."""
    vol = synth.synth_volume(synth.V_SPH, (80, 80, 80), orc.U8, 0x5EED)
    keys = all_keys(vol, 20, 2)
    script = [("vis", orc.RM_1DTRANS, (256.0, 0.0)), ("first",), ("vis", orc.RM_1DTRANS, (0.0, 255.0)),
              ("upload", keys[:7]), ("vis", orc.RM_ISOSURFACE, (90.0,)), ("upload", keys[5:30]),
              ("vis", orc.RM_2DTRANS, (10.0, 40.0, 0.0, 255.0)), ("upload", keys[40:44]),
              ("vis", orc.RM_1DTRANS, (256.0, 0.0))]
    replay(tmp_path, vol, orc.U8, 20, 2, (80, 60, 40), script)


def test_more_requests_than_slots(tmp_path):
    """UploadBrick refuses once every slot but the last was replaced in this call (GLVolumePool.cpp:771-772)."""
    vol = synth.synth_volume(synth.V_NOISE, (64, 64, 64), orc.U16, 7)
    keys = all_keys(vol, 20, 2)
    script = [("first",), ("vis", orc.RM_1DTRANS, (0.0, 65535.0)), ("upload", keys[:20]), ("upload", keys[20:45]),
              ("upload", keys[:3])]
    pool = replay(tmp_path, vol, orc.U16, 20, 2, (40, 40, 40), script)   # 8 slots
    assert pool.capacity == (2, 2, 2)


def test_metadata_texture_folding(tmp_path):
    """Fit1DIndexTo3DArray with a small GL_MAX_3D_TEXTURE_SIZE: the table folds to 2D / 3D."""
    vol = synth.synth_volume(synth.V_SPH, (96, 96, 96), orc.U8, 3)
    keys = all_keys(vol, 12, 2)
    for max3d in (64, 13):
        script = [("first",), ("vis", orc.RM_1DTRANS, (30.0, 200.0)), ("upload", keys[100:140])]
        pool = replay(tmp_path, vol, orc.U8, 12, 2, (48, 48, 48), script, max3d=max3d)
        assert pool.meta_dim[1] > 1


AUTOPOOL_CASES = [   # (volume x, y, z), dtype, brick, usable GPU memory in bytes, GL_MAX_3D_TEXTURE_SIZE
    ((44, 36, 28), orc.U8, 16, 1 << 20, 16384),          # budget below the dataset's need: the GPU layout wins
    ((44, 36, 28), orc.U8, 16, 64 << 20, 16384),         # everything fits: the dataset layout wins
    ((70, 45, 58), orc.U16, 20, 3 << 20, 16384),
    ((70, 45, 58), orc.U16, 20, 200 << 20, 16384),
    ((64, 64, 64), orc.U16, 36, 40 << 20, 16384),
    ((64, 64, 64), orc.U16, 36, 40 << 20, 128),          # the 3D texture limit clamps every axis
    ((96, 80, 40), orc.F32, 12, 5 << 20, 16384),
    ((96, 80, 40), orc.F32, 12, 5 << 20, 60),
]


@pytest.mark.parametrize("size,dtype,brick,budget,max3d", AUTOPOOL_CASES)
def test_pool_size_matches_reference_gpumemman(tmp_path, size, dtype, brick, budget, max3d):
    """the pool size the renderer picks when the host gives none -- orc.pool_size (restated in the product's
    size_pool) against the UNMODIFIED GPUMemMan::GetVolumePool (Renderer/GPUMemMan/GPUMemMan.cpp:766-844, compiled in place
    into oracle/_ref/ref_pool, directive `autopool`) over the reference's own dataset classes."""
    import subprocess
    vol = synth.synth_volume(synth.V_NOISE, size, dtype, 0x5EED)
    o = orc.Octree(vol, brick, 2)
    bits = {orc.U8: 8, orc.U16: 16, orc.F32: 32}[dtype]
    keys = list(o.iter_bricks())
    np.array([o.brick_size(*k) for k in keys], np.uint32).tofile(str(tmp_path / "sizes.bin"))
    np.ascontiguousarray(o.minmax, np.float64).tofile(str(tmp_path / "minmax.bin"))
    lines = ["vol %d %d %d" % tuple(size), "brick %d" % brick, "overlap 2", "bits %d" % bits, "float %d" % int(dtype == orc.F32),
             "pool 0 0 0", "max3d %d" % max3d, "lods %d" % o.lod_count]
    lines += ["layout %d %d %d %d" % ((lod,) + tuple(o.brick_count(lod))) for lod in range(o.lod_count)]
    lines += ["sizes %s" % (tmp_path / "sizes.bin"), "minmax %s" % (tmp_path / "minmax.bin"), "autopool %d" % budget]
    (tmp_path / "scenario.txt").write_text("\n".join(lines) + "\n")
    subprocess.check_call([pool_ref.BIN, str(tmp_path / "scenario.txt"), str(tmp_path / "result.txt"), str(tmp_path / "atlas.bin")],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    row = [l.split() for l in open(tmp_path / "result.txt") if l.startswith("autopool")][0]
    ref = tuple(int(v) for v in row[1:4])
    # GetMaxUsedBrickSizes: the largest brick that occurs (a volume smaller than one brick has smaller "max" bricks)
    used = tuple(max(o.brick_size(*k)[a] for k in keys) for a in range(3))
    got = tuple(orc.pool_size(budget, bits, 1, used, o.total_bricks, max3d))
    assert got == ref, (got, ref)
