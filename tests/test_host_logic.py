"""Host-side logic that needs no GPU: view/projection/LOD-factor helpers of the C ABI against an
independent numpy restatement, the TF mirrors against the oracle, the seeded volume generator."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import scene
import tuvok_b200 as tb
from oracle import orc
from tuvok_b200 import _lib as L
from tuvok_b200 import synth


def compute_view(w, h, rot, tra, eye=(0, 0, 1.6), at=(0, 0, 0), up=(0, 1, 0), fov=50.0):
    p = L.RenderParams()
    rc = tb.lib().tvk_compute_view(C.byref(p), w, h, L.f32x16(*rot.reshape(-1)), L.f32x16(*tra.reshape(-1)),
                                   L.f32x3(*eye), L.f32x3(*at), L.f32x3(*up), fov, 0.01, 1000.0, 1.0)
    assert rc == 0
    return p


@pytest.mark.parametrize("w,h", [(512, 512), (1920, 1080), (96, 64)])
def test_compute_view_matches_reference_conventions(w, h):
    rot = (tb.rotation_y(33.0) @ tb.rotation_x(-12.0)).astype(np.float32)
    tra = tb.translation(0.1, -0.05, 0.3)
    p = compute_view(w, h, rot, tra)
    view = scene.look_at((0, 0, 1.6), (0, 0, 0), (0, 1, 0))
    proj = scene.perspective(50.0, np.float32(w) / np.float32(h), 0.01, 1000.0)
    mv = ((rot @ tra).astype(np.float32) @ view).astype(np.float32)
    np.testing.assert_allclose(np.array(p.model_view).reshape(4, 4), mv, rtol=0, atol=2e-7)
    np.testing.assert_array_equal(np.array(p.projection, np.float32).reshape(4, 4), proj)
    assert np.float32(p.lod_factor) == scene.lod_factor(50.0, h)
    assert (p.width, p.height) == (w, h)


def test_look_at_is_gl_standard():
    # Basics/Vectors.h:1250-1265 with the default camera: translation by -1.6 along z, row-vector storage
    p = compute_view(64, 64, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32))
    mv = np.array(p.model_view).reshape(4, 4)
    expect = np.eye(4, dtype=np.float32)
    expect[3, 2] = -1.6
    np.testing.assert_allclose(mv, expect, atol=1e-7)


def test_default_params_are_abstrrenderer_defaults():
    p = L.RenderParams()
    assert tb.lib().tvk_default_params(C.byref(p), 640, 480) == 0
    # AbstrRenderer.cpp:77-171
    assert p.mode == tb.RM_1DTRANS and p.lighting == 1 and p.sample_rate_modifier == 1.0
    assert tuple(p.ambient) == (1, 1, 1, np.float32(0.1)) and tuple(p.diffuse) == (1, 1, 1, 1)
    assert tuple(p.light_dir) == (0, 0, -1) and tuple(p.iso_color) == (0.5, 0.5, 0.5)
    assert tuple(p.clip_min) == (0, 0, 0) and tuple(p.clip_max) == (1, 1, 1)
    assert tuple(p.eye) == (0, 0, np.float32(1.6))


@pytest.mark.parametrize("n,c,g", [(256, 0.5, 0.5), (4096, 0.3, 0.4), (256, 0.05, 0.5), (1024, 0.9, 0.3)])
def test_tf1d_mirror_matches_oracle(n, c, g):
    t = tb.TransferFunction1D(n)
    t.SetStdFunction(c, g)
    ref = orc.tf1d_std(n, c, g)
    np.testing.assert_array_equal(t.color, ref)
    np.testing.assert_array_equal(t.GetByteArray(), orc.tf1d_bytes(ref))
    assert t.GetNonZeroLimits() == orc.tf1d_nonzero(ref)


def test_tf1d_bytes_truncate():
    t = tb.TransferFunction1D(4)
    t.Set(np.array([[0.999, 0.5, 1.5, -1.0]] * 4, np.float32))
    assert tuple(t.GetByteArray()[0]) == (254, 127, 255, 0)     # (unsigned char)(v*255): truncation, clamped
    assert t.GetNonZeroLimits() == (0, 3)


def test_tf2d_nonzero_limits():
    t = tb.TransferFunction2D.rectangle(w=64, h=32, x0=0.25, x1=0.75, y0=0.5, y1=1.0)
    assert t.GetNonZeroLimits() == orc.tf2d_nonzero(t.GetByteArray())
    assert t.GetNonZeroLimits() == (16, 47, 16, 31)
    empty = tb.TransferFunction2D(np.zeros((4, 8, 4), np.uint8))
    assert empty.GetNonZeroLimits() == (8, 0, 4, 0)


def test_synth_is_seeded_and_pinned():
    # golden digests: the GPU generator is compared against these same arrays in the -m gpu tests
    v = synth.synth_volume(synth.V_NOISE, (40, 36, 32), 1, seed=0x5EED)
    assert v.dtype == np.uint16 and v.shape == (32, 36, 40)
    assert hashlib.sha1(v.tobytes()).hexdigest() == hashlib.sha1(
        synth.synth_volume(synth.V_NOISE, (40, 36, 32), 1, seed=0x5EED).tobytes()).hexdigest()
    assert not np.array_equal(v, synth.synth_volume(synth.V_NOISE, (40, 36, 32), 1, seed=1))
    r = synth.synth_volume(synth.V_RAMP, (8, 8, 1), 1)
    np.testing.assert_array_equal(r[0], np.arange(64).reshape(8, 8))        # x + 8y: the rebricking.h ramp
    s = synth.synth_volume(synth.V_SPH, (32, 32, 32), 0)
    assert s[0, 0, 0] == 0 and s.max() > 100                                 # empty exterior, shells inside
    f = synth.synth_volume(synth.V_SPH, (16, 16, 16), 2)
    assert f.dtype == np.float32 and 0.0 <= f.min() and f.max() <= 1.0
