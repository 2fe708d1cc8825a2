"""Stereo rendering of the 3D view: per-eye view / projection (FLOATMATRIX4::BuildStereoLookAtAndProjection,
Basics/Vectors.h:1215-1248, as GLRenderer::ComputeViewAndProjection calls it) and the eye composition of
GLRenderer::EndFrame (GLRenderer.cpp:758-812; Compose-{Anaglyphs,Scanline,SBS,AF}-FS.glsl).

CPU: oracle matrices == the reference's own matrix class (oracle/_ref/ref_host), tvk_compute_stereo_view == oracle,
oracle composition == the reference's shader text executed per fragment.  -m gpu: a whole stereo frame (both eyes
through the GridLeaper traversal kernel + the composition kernel) against the oracle, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import glsl_ref
import golden_scenes
import tuvok_b200 as tb
from oracle import orc
from scene import image_diff
from tuvok_b200 import _lib as L

HAVE_REF_HOST = os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_host"))
VIEWS = [((0, 0, 1.6), 50.0, 1920, 1080, 1.0, 0.02), ((0.4, -0.3, 2.2), 35.0, 96, 64, 2.5, 0.07), ((0, 0, 1.6), 50.0, 64, 64, 1.0, 0.0)]


@pytest.mark.skipif(not HAVE_REF_HOST, reason="oracle/_ref/ref_host not built (reference tree absent)")
@pytest.mark.parametrize("eye,fov,w,h,focal,dist", VIEWS)
def test_oracle_stereo_matrices_match_reference_matrix_class(tmp_path, eye, fov, w, h, focal, dist):
    import test_host_ref as hr
    aspect = np.float32(w) / np.float32(h)
    rows = hr.run(tmp_path, ["stereo %s 0 0 0 0 1 0 %.9g %.9g 0.01 1000 %.9g %.9g" % (hr.fl(eye), fov, aspect, focal, dist)])
    ref = [hr.hexf(r[1:]).reshape(4, 4) for r in rows]
    got = orc.stereo_view(eye, (0, 0, 0), (0, 1, 0), fov, float(aspect), 0.01, 1000.0, focal, dist)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("eye,fov,w,h,focal,dist", VIEWS)
def test_host_stereo_view_matches_oracle(eye, fov, w, h, focal, dist):
    from tuvok_b200.renderer import matmul4
    lib = tb.lib()
    rot = (tb.rotation_y(33.0) @ tb.rotation_x(-12.0)).astype(np.float32)
    tra = tb.translation(0.1, -0.05, 0.3)
    left, right = L.RenderParams(), L.RenderParams()
    rc = lib.tvk_compute_stereo_view(C.byref(left), C.byref(right), w, h, L.f32x16(*rot.reshape(-1)), L.f32x16(*tra.reshape(-1)),
                                     L.f32x3(*eye), L.f32x3(0, 0, 0), L.f32x3(0, 1, 0), fov, 0.01, 1000.0, 1.0, focal, dist)
    assert rc == 0
    vl, vr, pl, pr = orc.stereo_view(eye, (0, 0, 0), (0, 1, 0), fov, float(np.float32(w) / np.float32(h)), 0.01, 1000.0, focal, dist)
    rt = matmul4(rot, tra)
    for prm, v, p in ((left, vl, pl), (right, vr, pr)):
        assert np.array_equal(np.array(list(prm.projection), np.float32).reshape(4, 4), p)
        assert np.array_equal(np.array(list(prm.model_view), np.float32).reshape(4, 4), matmul4(rt, v))
        assert (prm.width, prm.height) == (w, h) and tuple(prm.eye) == tuple(np.float32(eye))
    assert left.lod_factor == right.lod_factor > 0
    if dist == 0.0:                                                  # no eye separation: both eyes are the mono camera
        mono = L.RenderParams()
        lib.tvk_compute_view(C.byref(mono), w, h, L.f32x16(*rot.reshape(-1)), L.f32x16(*tra.reshape(-1)), L.f32x3(*eye),
                             L.f32x3(0, 0, 0), L.f32x3(0, 1, 0), fov, 0.01, 1000.0, 1.0)
        assert list(mono.model_view) == list(left.model_view) == list(right.model_view)
        assert np.allclose(list(mono.projection), list(left.projection), rtol=1e-6, atol=0)


@pytest.mark.skipif(not glsl_ref.available(), reason="reference shaders absent")
def test_oracle_composition_matches_executed_reference_shaders(tmp_path):
    exe = glsl_ref.build_stereo(tmp_path)
    rng = np.random.default_rng(7)
    for (h, w) in [(48, 64), (37, 53), (1, 2)]:
        left, right = rng.random((h, w, 4), dtype=np.float32), rng.random((h, w, 4), dtype=np.float32)
        for mode in (orc.SM_RB, orc.SM_SCANLINE, orc.SM_SBS, orc.SM_AF):
            for swap in (False, True):
                for alt, split in ((0, 0.5), (1, 0.25)):
                    a = orc.stereo_compose(mode, left, right, swap, alt, split)
                    b = glsl_ref.run_stereo(exe, tmp_path, mode, left, right, swap, alt, split)
                    assert np.array_equal(a, b), (h, w, mode, swap, alt, split)


def test_composition_properties():
    rng = np.random.default_rng(3)
    h, w = 40, 64
    left, right = rng.random((h, w, 4), dtype=np.float32), rng.random((h, w, 4), dtype=np.float32)
    sc = orc.stereo_compose(orc.SM_SCANLINE, left, right)
    assert np.array_equal(sc[0::2], left[0::2]) and np.array_equal(sc[1::2], right[1::2])
    sbs = orc.stereo_compose(orc.SM_SBS, left, right)
    assert np.array_equal(sbs[:, :w // 2], left[:, 1::2]) and np.array_equal(sbs[:, w // 2:], right[:, 1::2])
    assert np.array_equal(orc.stereo_compose(orc.SM_AF, left, right, alternating_frame_id=0), left)
    assert np.array_equal(orc.stereo_compose(orc.SM_AF, left, right, alternating_frame_id=1), right)
    assert np.array_equal(orc.stereo_compose(orc.SM_AF, left, right, eye_swap=True), right)
    rb = orc.stereo_compose(orc.SM_RB, left, right)
    assert np.array_equal(rb[..., 3], np.maximum(left[..., 3], right[..., 3]))
    assert np.array_equal(rb[..., 1], rb[..., 2] * np.float32(0.5))
    assert np.allclose(rb[..., 0], left[..., :3] @ np.array([0.3, 0.59, 0.11], np.float32), atol=1e-6)


STEREO_SCENES = [("c2_bricked36_1d_ert", orc.SM_RB, False, 0.02), ("c3_bricked36_2d_lit", orc.SM_SCANLINE, False, 0.05),
                 ("ragged_1d_lit", orc.SM_SBS, True, 0.02), ("inside_aniso_2d", orc.SM_AF, False, 0.03)]


def oracle_stereo_frame(name, mode, swap, dist, focal=1.0, alt=0):
    """Both eyes through ONE pool, left first (GLGridLeaper renders the eyes of a frame through the same GLVolumePool):
    the right eye continues on the page table the left eye left behind."""
    eyes = []
    for e in (0, 1):
        s = golden_scenes.make(name)
        s.stereo_eye = (e, focal, dist)
        eyes.append(s.oracle_render(warm=eyes[0] if e == 1 else None))
    return eyes, orc.stereo_compose(mode, eyes[0]["image"], eyes[1]["image"], swap, alt, 0.5)


def test_oracle_eyes_are_horizontally_displaced():
    (l, r), img = oracle_stereo_frame("c2_bricked36_1d_ert", orc.SM_RB, False, 0.05)
    assert not np.array_equal(l["image"], r["image"])
    xs = np.arange(l["image"].shape[1], dtype=np.float64)

    def centroid(im):
        a = im[..., 3].astype(np.float64)
        return float((a * xs).sum() / a.sum())
    # the left view is translated by +eyeDist and its frustum shifted by +eyeDist * near / focalLength: with the
    # focal plane at 1.0 in front of a volume centred 1.6 away, the volume lies behind the plane of zero parallax
    # and appears further LEFT in the left eye
    assert centroid(l["image"]) < centroid(r["image"]) - 0.5
    assert img.shape == l["image"].shape


@pytest.mark.gpu
@pytest.mark.parametrize("name,mode,swap,dist", STEREO_SCENES)
def test_cuda_stereo_frame_matches_oracle(name, mode, swap, dist):
    (l, r), ref = oracle_stereo_frame(name, mode, swap, dist)
    s = golden_scenes.make(name)
    rd = s.make_renderer("device")
    rd.SetStereo(True)
    rd.SetStereoMode(mode)
    rd.SetStereoEyeSwap(swap)
    rd.SetStereoEyeDist(dist)
    assert rd.PaintStereoEye(0).converged
    fl = rd.ReadRGBA32F()
    assert np.array_equal(fl, l["image"])                            # fresh pool, same paging history: identical floats
    assert rd.PaintStereoEye(1).converged
    fr = rd.ReadRGBA32F()
    assert np.array_equal(fr, r["image"])                            # continues on the left eye's pool, as the oracle does
    assert np.array_equal(rd.page_table(), r["meta"])
    rd.ComposeStereo()
    f = rd.ReadRGBA32F()
    assert np.array_equal(f, orc.stereo_compose(mode, fl, fr, swap, 0, 0.5))   # the composition kernel, bit for bit
    assert np.array_equal(rd.ReadRGBA8(), orc.rgba8(f))
    assert np.array_equal(f, ref)
    # a mono frame afterwards is the mono image again (the composed frame does not stick)
    assert rd.PaintUntilConverged().converged
    mono = s.oracle_render(warm=r)
    assert np.array_equal(rd.ReadRGBA32F(), mono["image"])
    mx, psnr = image_diff(rd.ReadRGBA8(), s.oracle_render()["rgba8"])   # ... and, up to the paging history, the cold one
    assert mx <= 2 and psnr >= 45.0, (mx, psnr)
    rd.Cleanup()


@pytest.mark.gpu
def test_cuda_stereo_alternating_frames_and_isosurface():
    name = "c4_f32_iso" if "c4_f32_iso" in golden_scenes.SCENES else None
    if name is None:
        pytest.skip("no isosurface golden scene")
    (l, r), _ = oracle_stereo_frame(name, orc.SM_AF, False, 0.04)
    s = golden_scenes.make(name)
    rd = s.make_renderer("device")
    rd.SetStereoMode(orc.SM_AF)
    rd.SetStereoEyeDist(0.04)
    rd.PaintStereoUntilConverged()
    assert image_diff(rd.ReadRGBA8(), l["rgba8"])[0] <= 1
    rd.ToggleStereoFrame()
    rd.PaintStereoUntilConverged()
    assert image_diff(rd.ReadRGBA8(), r["rgba8"])[0] <= 1
    rd.Cleanup()
