"""-m gpu: images, page tables and miss lists of the CUDA renderer (through the C ABI) against the CPU
oracle and the committed golden vectors.  Contract (BASELINE.json north_star): page table / visibility /
min-max bit-exact; RGBA8 max |delta| <= 2/255 per channel and PSNR >= 45 dB.  The arithmetic contract
 is tight enough that the float images are in fact identical; that is asserted too where no
libm-dependent function (powf for sample rates != 1) is involved."""
import os

import numpy as np
import pytest

import golden_scenes
import tuvok_b200 as tb
from oracle import orc
from scene import Scene, image_diff
from tuvok_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_DELTA, MIN_PSNR = 2, 45.0   # the tolerance BASELINE.json states


def check_images(img8, ref8):
    mx, psnr = image_diff(img8, ref8)
    assert mx <= MAX_DELTA and psnr >= MIN_PSNR, "max|d| = %d/255, PSNR = %.1f dB" % (mx, psnr)
    return mx, psnr


@pytest.mark.parametrize("name", sorted(golden_scenes.SCENES))
@pytest.mark.parametrize("source", ["device", "callback"])
def test_golden_scene(name, source):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    s = golden_scenes.make(name)
    r = s.make_renderer(source)
    r.enable_counters(True)
    st = r.PaintUntilConverged()
    assert st.converged
    assert np.array_equal(r.page_table(), g["meta"])                 # page-table contents bit-exact
    if source == "device":
        assert np.array_equal(r.minmax(len(g["minmax"])), g["minmax"])   # brick min/max bit-exact
    img = r.ReadRGBA8()
    check_images(img, g["rgba8"])
    assert np.array_equal(img, g["rgba8"])
    assert np.array_equal(r.ReadRGBA32F(), g["image"])               # arithmetic contract: identical floats
    assert st.samples == int(g["samples"])
    r.Cleanup()


MODES = [(orc.RM_1DTRANS, False), (orc.RM_1DTRANS, True), (orc.RM_2DTRANS, False), (orc.RM_2DTRANS, True),
         (orc.RM_ISOSURFACE, True)]


@pytest.mark.parametrize("mode,lit", MODES)
@pytest.mark.parametrize("dtype", [orc.U8, orc.U16, orc.F32])
def test_all_modes_and_dtypes_against_live_oracle(mode, lit, dtype):
    iso = {orc.U8: 80, orc.U16: 20000, orc.F32: 0.3}[dtype]
    s = Scene(kind=synth.V_NOISE, size=(72, 64, 56), dtype=dtype, brick=20, overlap=2, width=112, height=80,
              mode=mode, lighting=lit, rotation=golden_scenes.ROT, tf_center=0.3, tf_inv_gradient=0.3, isovalue=iso)
    ref = s.oracle_render()
    r = s.make_renderer("device")
    r.enable_counters(True)
    st = r.PaintUntilConverged()
    assert st.converged and st.bricks_paged == ref["paged"]
    assert np.array_equal(r.page_table(), ref["meta"])
    check_images(r.ReadRGBA8(), ref["rgba8"])
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    assert st.samples == ref["stats"].samples and st.brick_visits == ref["stats"].brick_visits
    if mode == orc.RM_ISOSURFACE:      # the MRTs the compose pass reads
        hp, hn = r.ReadIsoBuffers()
        assert np.array_equal(hp.reshape(-1, 4), ref["outs"][0]) and np.array_equal(hn.reshape(-1, 4), ref["outs"][1])
    r.Cleanup()


def test_baked_36_brick_kernel_2d_lit_and_iso():
    for mode, lit, dtype, iso in [(orc.RM_2DTRANS, True, orc.U16, 0), (orc.RM_ISOSURFACE, True, orc.F32, 0.3),
                                  (orc.RM_1DTRANS, True, orc.U8, 0)]:
        s = Scene(kind=synth.V_NOISE, size=(128, 100, 90), dtype=dtype, brick=36, overlap=2, width=128, height=96,
                  mode=mode, lighting=lit, rotation=golden_scenes.ROT, isovalue=iso, tf_center=0.25, tf_inv_gradient=0.3)
        ref = s.oracle_render()
        r = s.make_renderer("device")
        assert r.PaintUntilConverged().converged
        assert np.array_equal(r.page_table(), ref["meta"])
        check_images(r.ReadRGBA8(), ref["rgba8"])
        assert np.array_equal(r.ReadRGBA32F(), ref["image"])
        r.Cleanup()


def test_sample_rate_modifier_and_opacity_correction():
    # powf differs between libm and CUDA by an ulp or two: tolerance only
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U16, brick=20, overlap=2, width=96, height=96,
              lighting=True, rotation=golden_scenes.ROT, sample_rate=2.0, tf_center=0.3, tf_inv_gradient=0.3)
    ref = s.oracle_render()
    r = s.make_renderer("device")
    assert r.PaintUntilConverged().converged
    check_images(r.ReadRGBA8(), ref["rgba8"])
    assert float(np.abs(r.ReadRGBA32F() - ref["image"]).max()) < 1e-5
    r.Cleanup()


def test_nearest_interpolant_and_one_voxel_ghost():
    for kw in (dict(nearest=True, lighting=True), dict(overlap=1, brick=18, lighting=True),
               dict(overlap=1, brick=18, mode=orc.RM_2DTRANS)):
        base = dict(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, width=96, height=96,
                    rotation=golden_scenes.ROT, tf_center=0.3, tf_inv_gradient=0.3)
        base.update(kw)
        s = Scene(**base)
        ref = s.oracle_render()
        r = s.make_renderer("device")
        assert r.PaintUntilConverged().converged
        assert np.array_equal(r.page_table(), ref["meta"])
        # with a 1-voxel ghost the gradient taps of the reference bleed into the neighbouring atlas slot;
        # the slot-linear pool reproduces that by addressing taps in the virtual atlas (k_raycast.cu Foot<!FAST>)
        check_images(r.ReadRGBA8(), ref["rgba8"])
        assert np.array_equal(r.ReadRGBA32F(), ref["image"])
        r.Cleanup()


def test_miss_reports_per_subframe_match_oracle():
    s = golden_scenes.make("ragged_1d_lit")
    ref = s.oracle_render()
    r = s.make_renderer("device")
    for k, want in enumerate(ref["requests"]):
        st = r.Paint()
        got = r.missing_list()
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist())), "subframe %d" % k
        assert np.array_equal(got, want)       # collision-free table: even the order (= slot assignment) agrees
        assert st.missing_reported == len(want)
    assert st.converged and not r.CheckForRedraw()
    r.Cleanup()


def test_default_509_entry_hash_table_converges_to_the_same_image():
    # with the reference's default table the per-subframe request SETS are schedule
    # dependent; the converged image and the set of resident bricks are not.
    s = golden_scenes.make("c2_bricked36_1d_ert")
    ref = s.oracle_render()
    s2 = golden_scenes.make("c2_bricked36_1d_ert", hash_size=13)     # tiny table: many subframes
    r = s2.make_renderer("device")
    st = r.PaintUntilConverged(max_subframes=200)
    assert st.converged
    res = lambda m: set(np.nonzero(m >= orc.BI_FLAG_COUNT)[0].tolist())
    assert res(r.page_table()) == res(ref["meta"])
    check_images(r.ReadRGBA8(), ref["rgba8"])
    # a different paging history restarts rays at different resume points (GLGridLeaper-blend.glsl:130-137):
    # direction, t and hence the depth fed to ComputeLOD are re-derived from resumePos, so the converged
    # floats agree closely (well below one 8-bit step) but not bit for bit -- in the reference as well
    assert float(np.abs(r.ReadRGBA32F() - ref["image"]).max()) < 1.0 / 255.0
    r.Cleanup()


def test_small_pool_evicts_and_still_matches():
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, width=96, height=96,
              rotation=golden_scenes.ROT, pool_size=(60, 60, 40), tf_center=0.3, tf_inv_gradient=0.3)
    ref = s.oracle_render()
    r = s.make_renderer("device")
    st = r.PaintUntilConverged(max_subframes=64)
    assert st.bricks_paged == ref["paged"]
    assert np.array_equal(r.page_table(), ref["meta"])
    check_images(r.ReadRGBA8(), ref["rgba8"])
    r.Cleanup()


def test_c5_shape_big_bricks_streamed_from_the_host():
    """BASELINE configs[4] in miniature: u8 volume in 128^3 bricks (inner 124, 2 MiB each) that never sit on the
    device as a whole -- the host source (Dataset::GetBrick stand-in) is asked brick by brick through the pinned
    double-buffered staging path, and the pool holds only 6 of the 21 bricks, so the frame converges over 5 paging
    subframes with eviction (12 bricks paged).  Page table, paging count and the float image equal the oracle's."""
    s = Scene(kind=synth.V_NOISE, size=(250, 200, 180), dtype=orc.U8, brick=128, overlap=2, width=320, height=400,
              rotation=golden_scenes.ROT, translation=tb.translation(0, 0, 0.3), pool_size=(384, 256, 128),
              tf_center=0.25, tf_inv_gradient=0.3)
    assert s.octree.total_bricks == 21
    ref = s.oracle_render(max_subframes=128)
    assert ref["paged"] > 6 and ref["subframes"] >= 4
    asked = []
    o = s.octree
    r = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)

    def get_brick(x, y, z, lod):
        asked.append((x, y, z, lod))
        return o.brick(x, y, z, lod)

    r.RegisterDataset(s.size, s.brick, s.overlap, s.dtype, o.minmax, get_brick, scale=s.scale,
                      max_gradient_magnitude=s.max_grad)
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
    r.Resize(s.width, s.height); r.SetRotation(s.rotation); r.SetTranslation(s.translation)
    r.CreateVolumePool(s._pool_size)
    assert tuple(r.info().pool_capacity) == (3, 2, 1)
    st = r.PaintUntilConverged(max_subframes=128)
    assert st.converged and st.bricks_paged == ref["paged"] and len(asked) == st.bricks_paged + 1   # + UploadFirstBrick
    assert np.array_equal(r.page_table(), ref["meta"])
    check_images(r.ReadRGBA8(), ref["rgba8"])
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    r.Cleanup()


def test_view_change_reuses_resident_bricks_and_is_deterministic():
    s = golden_scenes.make("c3_bricked36_2d_lit")
    r = s.make_renderer("device")
    assert r.PaintUntilConverged().converged
    a = r.ReadRGBA32F()
    r.SetRotation(tb.rotation_y(75.0))
    st = r.PaintUntilConverged()
    assert st.converged
    r.SetRotation(s.rotation)
    st = r.PaintUntilConverged()
    assert st.converged and st.bricks_paged == 0          # everything still resident
    b = r.ReadRGBA32F()
    assert float(np.abs(b - a).max()) < 0.5 / 255.0       # single pass now vs resumed passes before (see above)
    assert not r.CheckForRedraw()
    st = r.Paint()                                        # same state, same pass structure: bit-identical
    assert np.array_equal(r.ReadRGBA32F(), b)
    r.Cleanup()


def test_error_reporting_mirrors_reference_conventions():
    r = tb.CudaGridLeaper()
    with pytest.raises(tb.TvkError) as e:
        r.Paint()                                         # no dataset: T_ERROR + false in the reference
    assert "dataset" in str(e.value)
    with pytest.raises(tb.TvkError):
        r.BuildVolume(np.zeros((8, 8, 8), np.uint8), 4, 2)        # brick <= 2*overlap
    s = golden_scenes.make("c1_single_brick_1d")
    r2 = s.make_renderer("callback")
    r2.SetRendermode(7)
    with pytest.raises(tb.TvkError) as e:
        r2.Paint()
    assert "rendering mode" in str(e.value)
    r.Cleanup(); r2.Cleanup()


def test_brick_source_failure_is_reported():
    s = golden_scenes.make("ragged_1d_lit")
    o = s.octree
    calls = {"n": 0}

    def flaky(x, y, z, lod):
        calls["n"] += 1
        if calls["n"] > 3:
            raise IOError("disk gone")
        return o.brick(x, y, z, lod)

    r = tb.CudaGridLeaper(hash_table_size=s.hash_size())
    r.RegisterDataset(s.size, s.brick, s.overlap, s.dtype, o.minmax, flaky)
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetUseLighting(True); r.Resize(s.width, s.height)
    r.SetRotation(s.rotation)
    r.CreateVolumePool()
    with pytest.raises(tb.TvkError) as e:
        r.PaintUntilConverged()
    assert e.value.code == 5          # TVK_ERR_SOURCE
    r.Cleanup()


def test_c2_full_size_properties():
    """BASELINE configs[1] at full size (512^3 u16, 36^3 bricks, 1D TF + ERT, 1024x1024): the oracle would
    need minutes, so size-independent properties: determinism, convergence, idempotence of the paging loop,
    alpha bounded by early ray termination, uncovered pixels exactly zero."""
    import torch
    from tuvok_b200 import workloads
    w = workloads.WORKLOADS["c2"]
    n = w["size"][0]
    raw = torch.empty(n * n * n * 2, dtype=torch.uint8, device="cuda")
    r = tb.CudaGridLeaper(max_gpu_mem=8 << 30, hash_table_size=16 * 16 * 16 * 5 + 8)
    r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
    r.BuildVolume(raw.data_ptr(), w["brick"], w["overlap"], size=w["size"], dtype=w["dtype"])
    del raw
    t1, t2 = workloads.transfer_functions(w)
    r.Set1DTrans(t1); r.Set2DTrans(t2); r.SetRendermode(w["mode"]); r.SetUseLighting(w["lighting"])
    r.Resize(w["width"], w["height"]); r.SetRotation(workloads.orbit_rotation(3))
    r.CreateVolumePool()
    st = r.PaintUntilConverged()
    assert st.converged and st.bricks_paged > 100
    a = r.ReadRGBA32F()
    table = r.page_table()
    st2 = r.Paint()
    assert st2.converged and st2.bricks_paged == 0 and np.array_equal(r.page_table(), table)
    assert np.array_equal(r.ReadRGBA32F(), a)
    assert a[..., 3].max() <= 1.0 + 1e-6 and a.min() >= 0.0
    assert (a[..., 3] > 0.99).any()                       # early ray termination is active in this config
    assert not a[0, 0].any() and not a[-1, -1].any()      # image corners see no volume
    img = r.ReadRGBA8()
    assert np.array_equal(img, orc.rgba8(a))              # read-back conversion
    r.Cleanup()
