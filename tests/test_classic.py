"""Classic per-brick raycaster (GLRaycaster): CPU checks of the oracle's frame planning (LoD choice, brick list,
min/max culling, depth order) and -m gpu parity of the CUDA path (tvk_render_classic) with the oracle."""
import numpy as np
import pytest

import golden_scenes
import tuvok_b200 as tb
from oracle import orc
from scene import Scene, image_diff
from tuvok_b200 import synth

CLASSIC_SCENES = {
    "c1_single_brick_1d": {},
    "c2_bricked36_1d_ert": {},
    "ragged_1d_lit": {},
    "inside_aniso_2d": {},
    "c3_bricked36_2d_lit": {},
    # zoomed out: the planner picks a coarser LoD, fStepScale = 2 (powf opacity correction)
    "c2_far_lod1": dict(base="c2_bricked36_1d_ert", translation=tb.translation(0.0, 0.0, -2.2)),
    # 2D TF + lighting on ragged u8 bricks, off-axis
    "ragged_2d_lit": dict(base="ragged_1d_lit", mode=orc.RM_2DTRANS, lighting=True),
    # isosurface mode (GLRaycaster-ISO-FS + RefineIsosurface, nearest hit of all bricks by the depth test, Compose-FS)
    "iso_f32": dict(base="c4_f32_iso"),
    "iso_u16_noise": dict(base="c2_bricked36_1d_ert", mode=orc.RM_ISOSURFACE),
    "iso_u8_ragged": dict(base="ragged_1d_lit", mode=orc.RM_ISOSURFACE, isovalue=90.0),
    "iso_inside_aniso": dict(base="inside_aniso_2d", mode=orc.RM_ISOSURFACE),
}


def make(name):
    kw = dict(CLASSIC_SCENES[name])
    base = kw.pop("base", name)
    return golden_scenes.make(base, **kw)


def test_lod_follows_screen_space_error():
    near = make("c2_bricked36_1d_ert").oracle_classic()
    far = make("c2_far_lod1").oracle_classic()
    assert near["lod"] == 0 and far["lod"] >= 1


@pytest.mark.parametrize("name", ["c2_bricked36_1d_ert", "ragged_1d_lit", "inside_aniso_2d"])
def test_brick_list_is_culled_and_depth_sorted(name):
    s = make(name)
    r = s.oracle_classic()
    o = s.octree
    bc = o.brick_count(r["lod"])
    n_lod = bc[0] * bc[1] * bc[2]
    idx, empty = r["order"][:, 0], r["order"][:, 1]
    assert len(set(idx.tolist())) == len(idx) <= n_lod            # every brick at most once
    d = r["distance"][empty == 0]
    assert (np.diff(d) >= 0).all()                                 # front to back
    # bIsEmpty is the legacy one-sided/two-sided min/max test of UVFDataset::ContainsData
    first = o.brick_index(0, 0, 0, r["lod"])
    lo, hi = s.visibility_args()[:2]
    for i, e in zip(idx, empty):
        mn, mx = o.minmax[first + i][:2]
        if s.mode == orc.RM_1DTRANS:
            assert bool(e) == (not (hi >= mn and lo <= mx))


def test_frustum_culling_drops_bricks_outside_the_view():
    s = golden_scenes.make("c2_bricked36_1d_ert", translation=tb.translation(0.9, 0.0, 1.0))   # volume half off screen
    r = s.oracle_classic()
    bc = s.octree.brick_count(r["lod"])
    assert 0 < len(r["order"]) < bc[0] * bc[1] * bc[2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CLASSIC_SCENES))
def test_cuda_classic_matches_oracle(name):
    s = make(name)
    ref = s.oracle_classic()
    r = s.make_renderer("device")
    r.enable_counters(True)
    st = r.PaintClassic()
    lod, order, dist = r.classic_brick_list()
    assert lod == ref["lod"]
    assert np.array_equal(order, ref["order"])                     # visibility list: same bricks, same flags, same order
    assert np.array_equal(dist, ref["distance"])
    img = r.ReadRGBA8()
    mx, psnr = image_diff(img, ref["rgba8"])
    assert mx <= 2 and psnr >= 45.0, (mx, psnr)
    f = r.ReadRGBA32F()
    if s.mode == orc.RM_ISOSURFACE:
        hp, hn = r.ReadIsoBuffers()
        assert np.array_equal(hp.reshape(-1, 4), ref["hit_pos"])    # hit position + fInterpolParam of the winning brick
        assert np.array_equal(hn.reshape(-1, 4), ref["hit_normal"]) # normal + iTileID
        assert np.array_equal(f, ref["image"])
        assert 0 < st.samples <= ref["samples"]                    # bricks behind the kept hit are not marched
    elif ref["lod"] == 0:                                          # no powf (libm vs CUDA) involved: identical floats
        assert np.array_equal(f, ref["image"])
        assert st.samples == ref["samples"]
    else:
        assert float(np.abs(f - ref["image"]).max()) < 1e-4
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_classic_callback_source_and_gridleaper_coexist():
    # the classic path pages through the same pool as GridLeaper; both can be used on one renderer
    s = make("ragged_1d_lit")
    ref_c, ref_g = s.oracle_classic(), s.oracle_render()
    r = s.make_renderer("callback")
    assert r.PaintUntilConverged().converged
    assert image_diff(r.ReadRGBA8(), ref_g["rgba8"])[0] <= 2
    r.PaintClassic()
    assert np.array_equal(r.ReadRGBA32F(), ref_c["image"])
    assert r.PaintUntilConverged().converged
    assert image_diff(r.ReadRGBA8(), ref_g["rgba8"])[0] <= 2
    r.Cleanup()


def test_classic_isosurface_agrees_with_the_gridleaper_isosurface():
    """Two different traversals of the same surface (per-brick first hit under the depth test vs the page-table walk):
    hit masks agree up to silhouette pixels and the shaded images are close."""
    s = make("iso_f32")
    c, g = s.oracle_classic(), s.oracle_render()
    hc, hg = c["hit_pos"][:, 3] != 0, g["outs"][0].reshape(-1, 4)[:, 3] != 0
    assert hc.sum() > 100 and float((hc != hg).mean()) < 0.01
    assert float(np.abs(c["image"] - g["image"]).max()) < 2.0 / 255.0
    # every hit carries the list position of a non-empty brick
    tiles = c["hit_normal"][hc, 3].astype(int)
    assert (c["order"][tiles, 1] == 0).all()


CLEARVIEW = [("c4_f32_iso", {}, dict(isovalue=0.8, size=2.5)),
             ("c2_bricked36_1d_ert", dict(mode=orc.RM_ISOSURFACE), dict(isovalue=45000.0, color=(0.2, 0.9, 0.1), context_scale=2.0)),
             ("ragged_1d_lit", dict(mode=orc.RM_ISOSURFACE, isovalue=90.0), dict(isovalue=140.0, size=3.0, border_scale=20.0,
                                                                              pos=(0.1, -0.1, 0.3, 1.0))),
             ("inside_aniso_2d", dict(mode=orc.RM_ISOSURFACE), dict(isovalue=50000.0))]


def test_clearview_focus_surface_lies_behind_the_context_surface():
    s = golden_scenes.make("c4_f32_iso")
    s.clearview = dict(isovalue=0.8, size=2.5)
    r = s.oracle_classic()
    ctx, foc = r["hit_pos"][:, 3] != 0, r["cv_pos"][:, 3] != 0
    assert foc.sum() > 0 and not (foc & ~ctx).any()                 # a focus hit needs iso_cv >= iso: the context is hit first
    both = ctx & foc
    assert (-r["cv_pos"][both, 2] >= -r["hit_pos"][both, 2] - 1e-4).all()   # ... and lies farther from the eye
    plain = golden_scenes.make("c4_f32_iso").oracle_classic()
    assert np.array_equal(plain["hit_pos"], r["hit_pos"])            # the first pass does not depend on ClearView
    assert not np.array_equal(plain["image"], r["image"])


@pytest.mark.gpu
@pytest.mark.parametrize("name,over,cv", CLEARVIEW)
def test_cuda_clearview_matches_oracle(name, over, cv):
    s = golden_scenes.make(name, **over)
    s.clearview = dict(cv)
    ref = s.oracle_classic()
    r = s.make_renderer("device")
    r.SetCV(True)
    r.SetCVIsoValue(cv["isovalue"])
    if "color" in cv: r.SetCVColor(cv["color"])
    if "size" in cv: r.SetCVSize(cv["size"])
    if "context_scale" in cv: r.SetCVContextScale(cv["context_scale"])
    if "border_scale" in cv: r.SetCVBorderScale(cv["border_scale"])
    if "pos" in cv: r.SetCVFocusPos(cv["pos"])
    r.PaintClassic()
    hp, hn = r.ReadIsoBuffers()
    cp, cn = r.ReadCVBuffers()
    assert np.array_equal(hp.reshape(-1, 4), ref["hit_pos"]) and np.array_equal(hn.reshape(-1, 4), ref["hit_normal"])
    assert np.array_equal(cp.reshape(-1, 4), ref["cv_pos"]) and np.array_equal(cn.reshape(-1, 4), ref["cv_normal"])
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    assert np.array_equal(r.ReadRGBA8(), ref["rgba8"])
    # ClearView off again: the plain isosurface frame
    r.SetCV(False)
    r.PaintClassic()
    plain = golden_scenes.make(name, **over).oracle_classic()
    assert np.array_equal(r.ReadRGBA32F(), plain["image"])
    r.Cleanup()
