"""HQ MIP frames of the 2D windows (GLRenderer.cpp:1183-1285, AbstrRenderer::PlanHQMIPFrame,
GLRaycaster::RenderHQMIPPreLoop / RenderHQMIPInLoop, GLRaycaster-MIP-Rot-FS.glsl, Transfer-MIP-FS.glsl).

CPU: the oracle restatement (orc_mip_*) against the reference's own shader text executed per brick (tests/glsl_ref.py),
the MIP rotation against the reference's FLOATMATRIX4 (oracle/_ref/ref_host), PlanHQMIPFrame's LoD rule by hand.
-m gpu: the CUDA path (tvk_render_mip) against the oracle, bit for bit."""
import os

import numpy as np
import pytest

import glsl_ref
import golden_scenes
import tuvok_b200 as tb
from oracle import orc
from scene import image_diff

# name -> (base scene, scene overrides, window mode, flip, MIP rotation angle, m_bMIPLOD)
MIP_SCENES = {
    "coronal_u16": ("c2_bricked36_1d_ert", {}, "coronal", (False, False), 0.0, True),
    "sagittal_rot_u8_ragged": ("ragged_1d_lit", {}, "sagittal", (False, False), 33.0, True),
    "axial_flip_aniso": ("inside_aniso_2d", dict(translation=None), "axial", (True, False), 200.0, True),
    "coronal_small_window_lod1": ("c2_bricked36_1d_ert", dict(width=40, height=36), "coronal", (False, True), 15.0, True),
    "coronal_small_window_lod0": ("c2_bricked36_1d_ert", dict(width=40, height=36), "coronal", (False, True), 15.0, False),
    "single_brick": ("c1_single_brick_1d", {}, "sagittal", (True, True), 77.0, True),
}


def make(name):
    base, over, wm, flip, angle, mip_lod = MIP_SCENES[name]
    kw = dict(over)
    kw["rotation"] = tb.mip_rotation(wm, angle, flip)
    kw.setdefault("translation", None)
    return golden_scenes.make(base, **kw), wm, flip, angle, mip_lod


def test_plan_hq_mip_frame_lod_rule():
    p = orc.RenderParams()
    for vol, win, count, use, want in [((96, 96, 96), (80, 80), 3, 1, 0),      # one halving -> stepped back to 0
                                       ((96, 96, 96), (40, 36), 3, 1, 1),      # 96 >= 40, 48 >= 40, 24 < 40 -> 2 - 1
                                       ((96, 96, 96), (40, 36), 3, 0, 0),      # m_bMIPLOD off
                                       ((2048, 2048, 2048), (1920, 1080), 7, 1, 0),
                                       ((2048, 2048, 512), (512, 512), 7, 1, 0),   # smallest extent decides: 512>=512, 256<512
                                       ((4096, 4096, 4096), (256, 256), 3, 1, 2),  # clamped to the last LoD
                                       ((16, 16, 16), (1, 1), 9, 1, 4)]:       # 16,8,4,2,1 >= 1, then 0 >= 1 fails: 5 - 1
        p.vol = (C_U32 * 3)(*vol)
        p.width, p.height = win
        assert orc.mip_lod(p, count, use) == want, (vol, win, count, use)


import ctypes
C_U32 = ctypes.c_uint32


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_host")),
                    reason="oracle/_ref/ref_host not built (reference tree absent)")
@pytest.mark.parametrize("wm,code", [("sagittal", 0), ("axial", 1), ("coronal", 2)])
@pytest.mark.parametrize("flip", [(False, False), (True, False), (True, True)])
def test_mip_rotation_matches_reference_matrix_class(tmp_path, wm, code, flip):
    import test_host_ref as hr
    reg = (tb.rotation_y(12.0) @ tb.rotation_x(-7.0)).astype(np.float32)
    view = np.eye(4, dtype=np.float32)
    view[3, :3] = (0.0, 0.0, -1.6)
    for angle in (0.0, 33.0, 200.0):
        rows = hr.run(tmp_path, ["miprot %d %d %d %.9g %s %s" % (code, flip[0], flip[1], angle, hr.fl(reg), hr.fl(view))])
        ref = hr.hexf(rows[0][1:]).reshape(4, 4)
        got = tb.mip_rotation(wm, angle, flip, reg)
        assert np.array_equal(got, ref), (wm, flip, angle, np.abs(got - ref).max())
        from tuvok_b200.renderer import matmul4
        assert np.array_equal(matmul4(got, view), hr.hexf(rows[1][1:]).reshape(4, 4))


@pytest.mark.skipif(not glsl_ref.available(), reason="reference shaders / oracle/_ref tools absent")
@pytest.mark.parametrize("name", sorted(MIP_SCENES))
def test_oracle_mip_matches_executed_reference_shaders(tmp_path, name):
    s, *_rest, mip_lod = make(name)
    r = s.oracle_mip(use_mip_lod=mip_lod)
    p = r["params"]
    u = orc.uniforms(p)
    exe = glsl_ref.build_mip(tmp_path)
    img, mx = glsl_ref.run_mip(exe, tmp_path, p, u["inv_proj"], u["mv_inv"], u["norm"], r["bricks"], r["n"], r["data"],
                               s.tf1d.GetByteArray())
    om = r["max"].reshape(-1, 2)
    assert np.array_equal(om[:, 1], mx[:, 3])                       # coverage: the same fragments were generated
    assert np.array_equal(mx[:, 0], mx[:, 1]) and np.array_equal(mx[:, 0], mx[:, 2])
    assert float(np.abs(om[:, 0] - mx[:, 0]).max()) <= 5e-5         # maxima to rounding (unfused shader text)
    a8, b8 = orc.rgba8(r["image"]), orc.rgba8(img.reshape(s.height, s.width, 4))
    d8, psnr = image_diff(a8, b8)
    assert d8 <= 2 and psnr >= 45.0
    assert float((a8 == b8).all(axis=2).mean()) >= 0.995            # a maximum on a TF bin edge may land one bin apart
    assert (img[:, 3] == 1.0).all() and (r["image"][..., 3] == 1.0).all()
    assert (r["image"].reshape(-1, 4)[om[:, 1] == 0, :3] == 0).all()   # uncovered pixels are black


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_host")),
                    reason="oracle/_ref/ref_host not built (reference tree absent)")
@pytest.mark.parametrize("w,h", [(96, 96), (128, 72), (72, 128), (1920, 1080), (333, 777), (40, 36)])
def test_mip_ortho_projection_matches_reference_statements(tmp_path, w, h):
    """GLRenderer.cpp:1183-1197 run on the reference's own DOUBLEVECTOR2 / FLOATMATRIX4::Ortho (ref_host mipo)."""
    import test_host_ref as hr
    rows = hr.run(tmp_path, ["mipo %d %d" % (w, h)])
    ref = hr.hexf(rows[0][1:]).reshape(4, 4)
    got = tb.mip_ortho_projection(w, h)
    assert np.array_equal(got, ref), np.abs(got - ref).max()
    assert got[2, 3] == 0.0 and got[3, 3] == 1.0                    # parallel: w' does not depend on z


ORTHO_SCENES = ["coronal_u16", "sagittal_rot_u8_ragged", "axial_flip_aniso", "coronal_small_window_lod1"]


def make_ortho(name):
    s, wm, flip, angle, mip_lod = make(name)
    s.ortho_mip = True                                              # AbstrRenderer::SetOrthoView(true)
    return s, wm, flip, angle, mip_lod


@pytest.mark.skipif(not glsl_ref.available(), reason="reference shaders / oracle/_ref tools absent")
@pytest.mark.parametrize("name", ORTHO_SCENES)
def test_oracle_ortho_mip_matches_executed_reference_shaders(tmp_path, name):
    """m_bOrthoView (GLRenderer.cpp:1183-1197, GLRaycaster.cpp:486-487): parallel rays, model view = the MIP rotation alone."""
    s, *_rest, mip_lod = make_ortho(name)
    r = s.oracle_mip(use_mip_lod=mip_lod)
    p = r["params"]
    u = orc.uniforms(p)
    assert u["inv_proj"][11] == 0.0
    exe = glsl_ref.build_mip(tmp_path)
    img, mx = glsl_ref.run_mip(exe, tmp_path, p, u["inv_proj"], u["mv_inv"], u["norm"], r["bricks"], r["n"], r["data"],
                               s.tf1d.GetByteArray())
    om = r["max"].reshape(-1, 2)
    assert 0.05 < om[:, 1].mean() < 1.0                             # the volume is on screen, with a margin around it
    assert np.array_equal(om[:, 1], mx[:, 3])
    assert float(np.abs(om[:, 0] - mx[:, 0]).max()) <= 5e-5
    a8, b8 = orc.rgba8(r["image"]), orc.rgba8(img.reshape(s.height, s.width, 4))
    d8, psnr = image_diff(a8, b8)
    assert d8 <= 2 and psnr >= 45.0
    # parallel rays: the covered region of an unrotated coronal view is an axis-aligned rectangle
    if name == "coronal_u16":
        cov = r["max"][..., 1] > 0
        ys, xs = np.nonzero(cov)
        assert cov[ys.min():ys.max() + 1, xs.min():xs.max() + 1].all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ORTHO_SCENES)
def test_cuda_ortho_mip_matches_oracle(name):
    s, wm, flip, angle, mip_lod = make_ortho(name)
    ref = s.oracle_mip(use_mip_lod=mip_lod)
    r = s.make_renderer("device")
    r.SetMIPRotationAngle(angle)
    r.SetMIPLOD(mip_lod)
    r.SetOrthoView(True)
    r.PaintHQMIP(wm, flip)
    lod, order, _ = r.classic_brick_list()
    assert lod == ref["lod"] and np.array_equal(order, ref["order"])
    assert np.array_equal(r.mip_max_image(), ref["max"])
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    # and back: the perspective frame of the same renderer is unchanged by the excursion
    r.SetOrthoView(False)
    s.ortho_mip = False
    r.PaintHQMIP(wm, flip)
    assert np.array_equal(r.mip_max_image(), s.oracle_mip(use_mip_lod=mip_lod)["max"])
    with pytest.raises(tb.TvkError):                                # a parallel projection outside the MIP frame is refused
        r.SetOrthoView(True)
        r._push_params(); r._push_ortho_mip(np.eye(4, dtype=np.float32)); r._ck(r._lib.tvk_render_classic(r._h, None))
    r.Cleanup()


def test_mip_properties():
    s, *_ = make("sagittal_rot_u8_ragged")
    r = s.oracle_mip()
    m = r["max"]
    assert 0 < m[..., 1].mean() < 1                                  # some pixels covered, some not
    # the maximum along a ray can never exceed the largest voxel of the (non-empty) bricks it crosses
    top = max(float(np.max(d)) for d in r["data"] if d is not None) * orc.uniforms(r["params"])["norm"]
    assert float(m[..., 0].max()) <= top + 1e-7
    # empty bricks are skipped (GLRenderer.cpp:1209-1211), every other brick of the LoD is listed: no frustum culling
    bc = s.octree.brick_count(r["lod"])
    assert r["n"] == bc[0] * bc[1] * bc[2]
    assert np.array_equal(r["order"][:, 0], np.arange(r["n"]))       # key order
    # LoD choice ignores the camera: a translated view lists the same bricks
    s2 = golden_scenes.make("ragged_1d_lit", rotation=s.rotation, translation=tb.translation(0.8, 0.0, 0.0))
    assert np.array_equal(s2.oracle_mip()["order"], r["order"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MIP_SCENES))
def test_cuda_mip_matches_oracle(name):
    s, wm, flip, angle, mip_lod = make(name)
    ref = s.oracle_mip(use_mip_lod=mip_lod)
    r = s.make_renderer("device")
    r.enable_counters(True)
    r.SetMIPRotationAngle(angle)
    r.SetMIPLOD(mip_lod)
    st = r.PaintHQMIP(wm, flip)
    lod, order, _ = r.classic_brick_list()
    assert lod == ref["lod"]
    assert np.array_equal(order, ref["order"])
    assert np.array_equal(r.mip_max_image(), ref["max"])             # blended maxima and coverage, bit for bit
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    assert np.array_equal(r.ReadRGBA8(), ref["rgba8"])
    # A ray that lies exactly in the plane two bricks share (axis-aligned MIP views produce a few) belongs to both
    # bricks for the oracle's per-brick slab test and to one cell for the kernel's grid walk; the boundary samples
    # are the same voxels (ghost layers), so the maxima above are identical and only the count may differ.
    assert abs(int(st.samples) - int(ref["samples"])) <= 1e-4 * ref["samples"]
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_mip_2d_mode_callback_source_and_coexistence():
    """Brick emptiness follows the current render mode (2D TF here), the colour the 1D TF; MIP frames, classic frames and
    GridLeaper frames share one pool."""
    s = golden_scenes.make("inside_aniso_2d", rotation=tb.mip_rotation("coronal", 45.0), translation=None)
    ref = s.oracle_mip()
    ref_g = s.oracle_render()
    r = s.make_renderer("callback")
    r.SetMIPRotationAngle(45.0)
    assert r.PaintUntilConverged().converged
    assert image_diff(r.ReadRGBA8(), ref_g["rgba8"])[0] <= 2
    r.PaintHQMIP("coronal")
    assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    assert r.PaintUntilConverged().converged
    assert image_diff(r.ReadRGBA8(), ref_g["rgba8"])[0] <= 2
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_mip_turntable_reuses_the_plan_and_follows_tf_changes():
    """A MIP turntable: the frame plan (brick list, per-brick tables) does not depend on the view and is kept across
    frames; a transfer-function change, a classic frame in between or a resize re-plan.  Every frame equals the oracle."""
    from tuvok_b200.tf import TransferFunction1D
    base = dict(width=72, height=60)
    r = None
    for step, (angle, tf_c) in enumerate([(0.0, 0.4), (33.0, 0.4), (66.0, 0.4), (66.0, 0.7), (120.0, 0.7)]):
        s = golden_scenes.make("c2_bricked36_1d_ert", rotation=tb.mip_rotation("sagittal", angle), translation=None,
                               tf_center=tf_c, **base)
        ref = s.oracle_mip()
        if r is None:
            r = s.make_renderer("device")
        else:
            r.Set1DTrans(s.tf1d)
        if step == 2:
            r.PaintClassic()                                          # overwrites the brick list: the next MIP frame re-plans
        r.SetMIPRotationAngle(angle)
        r.PaintHQMIP("sagittal")
        assert np.array_equal(r.classic_brick_list()[1], ref["order"])
        assert np.array_equal(r.mip_max_image(), ref["max"])
        assert np.array_equal(r.ReadRGBA32F(), ref["image"])
    r.Cleanup()
