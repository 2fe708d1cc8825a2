"""-m gpu: the procedural multi-resolution dataset (BASELINE configs[4]) through the C ABI -- min/max table evaluated on
the device == min/max of the generated bricks (numpy), bricks paged in == the numpy field, the rendered frame re-traced by
the CPU oracle on the renderer's own page table and pool (tests/parity_gate.py), host brick cache hits after eviction."""
import numpy as np
import pytest

import parity_gate
import tuvok_b200 as tb
from oracle import orc
from tuvok_b200 import _lib as L
from tuvok_b200 import synth
from tuvok_b200.tf import TransferFunction1D, TransferFunction2D

pytestmark = pytest.mark.gpu


def make(kind, size, dtype, brick, overlap, mode=L.RM_1DTRANS, lighting=True, w=160, h=120, cache=0, max_gpu_mem=1 << 30,
         minmax=None, pool=None):
    sizes, layouts, offs = synth.procedural_geometry(size, brick, overlap)
    r = tb.CudaGridLeaper(max_gpu_mem=max_gpu_mem, hash_table_size=offs[1] * 7 + 8)
    r.SetProceduralVolume(kind, size, dtype, brick, overlap, minmax=minmax, host_cache_bytes=cache, threads=4)
    n = 256 if dtype == L.U8 else 4096
    t1 = TransferFunction1D(n)
    t1.SetStdFunction(0.25, 0.3)
    r.Set1DTrans(t1)
    r.Set2DTrans(TransferFunction2D.rectangle(w=n, h=64, x0=0.02, x1=0.9, alpha_max=64))
    r.SetRendermode(mode)
    r.SetUseLighting(lighting)
    r.Resize(w, h)
    r.SetRotation((tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32))
    r.CreateVolumePool(pool)
    return r, (sizes, layouts, offs)


@pytest.mark.parametrize("kind,dtype", [(synth.V_NOISE, L.U8), (synth.V_SPH, L.U16), (synth.V_NOISE, L.F32)])
def test_minmax_table_and_bricks(kind, dtype):
    size, brick, ov = (150, 120, 100), 36, 2
    r, (sizes, layouts, offs) = make(kind, size, dtype, brick, ov)
    mm = r.minmax(offs[-1]).reshape(-1, 4)
    sl = r.procedural_minmax(kind, size, dtype, brick, ov, 5, 11)                  # a slice, as a rank of N would take it
    assert np.array_equal(sl, mm[5:16])
    rng = np.random.default_rng(3)
    for lod, lay in enumerate(layouts):
        for _ in range(6):
            x, y, z = (int(rng.integers(0, l)) for l in lay)
            ref = synth.procedural_brick(kind, size, dtype, 0x5EED, brick, ov, x, y, z, lod)
            i = offs[lod] + x + lay[0] * (y + lay[1] * z)
            assert mm[i, 0] == float(ref.min()) and mm[i, 1] == float(ref.max()), (lod, x, y, z)
            assert np.array_equal(r.brick(x, y, z, lod, dtype), ref)
    r.Cleanup()


@pytest.mark.parametrize("mode,lighting,dtype,brick", [(L.RM_1DTRANS, True, L.U8, 36), (L.RM_2DTRANS, True, L.U16, 36),
                                                       (L.RM_1DTRANS, False, L.U8, 68), (L.RM_ISOSURFACE, True, L.F32, 20)])
def test_frame_is_the_oracles(mode, lighting, dtype, brick):
    size, ov = (200, 160, 180), 2
    r, _ = make(synth.V_NOISE, size, dtype, brick, ov, mode=mode, lighting=lighting)
    if mode == L.RM_ISOSURFACE:
        r.SetIsoValue(0.2)
    res = parity_gate.check_frame(r, size, dtype, brick, ov, stride=2)
    assert res["ok"] and res["float_bit_identical"] and res["pixels"] > 500, res
    st = r.stream_stats()
    assert st.bricks_uploaded > 0 and st.bricks_generated == st.bricks_uploaded and st.host_cache_hits == 0
    r.Cleanup()


def test_supplied_minmax_table_gives_the_same_frame():
    size, brick, ov = (200, 160, 180), 36, 2
    a, (_, _, offs) = make(synth.V_NOISE, size, L.U8, brick, ov)
    mm = a.procedural_minmax(synth.V_NOISE, size, L.U8, brick, ov, 0, offs[-1])
    assert a.PaintUntilConverged().converged
    img = a.ReadRGBA32F().copy()
    a.Cleanup()
    # the table assembled from two slices (what two ranks would compute) handed in
    b0, _ = make(synth.V_NOISE, size, L.U8, brick, ov, minmax=mm)
    assert b0.PaintUntilConverged().converged
    assert np.array_equal(b0.ReadRGBA32F(), img)
    b0.Cleanup()


def test_host_cache_serves_bricks_the_pool_evicted():
    """a pool that holds one view's working set but not two: going back and forth pages bricks in again -- from the host
    cache, not from the generator -- and the frames stay what they were"""
    size, brick, ov = (260, 260, 260), 36, 2
    views = [(tb.rotation_y(a) @ tb.rotation_x(20.0)).astype(np.float32) for a in (0.0, 180.0)]
    need, frames = [], []
    for v in views:                                       # working set of each view alone, and its frame
        r, _ = make(synth.V_NOISE, size, L.U8, brick, ov, w=320, h=240)
        r.SetRotation(v)
        assert r.PaintUntilConverged().converged
        need.append(int(r.stream_stats().bricks_uploaded))
        frames.append(r.ReadRGBA8().copy())
        r.Cleanup()
    r, _ = make(synth.V_NOISE, size, L.U8, brick, ov, w=320, h=240)
    for v in views:
        r.SetRotation(v)
        assert r.PaintUntilConverged().converged
    union = int(r.stream_stats().bricks_uploaded)          # both working sets together (large pool: nothing evicted)
    r.Cleanup()
    assert union - max(need) >= 16, (need, union)
    slots = max(need) + (union - max(need)) // 2            # one view fits with room to spare, both do not
    r, _ = make(synth.V_NOISE, size, L.U8, brick, ov, w=320, h=240, cache=256 << 20, pool=(36 * slots, 36, 36))
    for k in range(2):
        r.SetRotation(views[k])
        assert r.PaintUntilConverged().converged
    gen0 = r.stream_stats().bricks_generated
    up0 = r.stream_stats().bricks_uploaded
    for k in range(4):
        r.SetRotation(views[k % 2])
        assert r.PaintUntilConverged().converged
        # against the frame of a fresh renderer only the reference's own history dependence remains (DESIGN.md section 4:
        # sample positions are slot-relative, and this transfer function is steep)
        mx, psnr = parity_gate.image_metrics(r.ReadRGBA8(), frames[k % 2])
        assert psnr >= 45.0, (mx, psnr)
    # the bricks that came back from the host cache hold the right voxels: resident slots == the numpy field
    sizes, layouts, offs = synth.procedural_geometry(size, brick, ov)
    meta = r.page_table()
    resident = np.flatnonzero(meta[:offs[-1]] >= 3)
    assert resident.size >= 40
    for i in resident[:: max(1, resident.size // 40)]:
        lod = int(np.searchsorted(offs, i, side="right") - 1)
        lay = layouts[lod]
        local = int(i) - offs[lod]
        x, y, z = local % lay[0], (local // lay[0]) % lay[1], local // (lay[0] * lay[1])
        ref = synth.procedural_brick(synth.V_NOISE, size, L.U8, 0x5EED, brick, ov, x, y, z, lod)
        got = r.pool_slot(int(meta[i]) - 3, L.U8, (brick, brick, brick))
        assert np.array_equal(got[:ref.shape[0], :ref.shape[1], :ref.shape[2]], ref), (lod, x, y, z)
    st = r.stream_stats()
    assert st.bricks_uploaded > up0                       # bricks were paged in again ...
    assert st.host_cache_hits > 0 and st.bricks_generated + st.host_cache_hits >= st.bricks_uploaded
    assert st.bricks_generated - gen0 < st.bricks_uploaded - up0   # ... and not all of them by generating them anew
    r.Cleanup()
