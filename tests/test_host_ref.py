"""Host-side arithmetic in front of the ray caster against UNMODIFIED reference code (oracle/_ref/ref_host,
compiled in place from /root/reference: Basics/Vectors.h, Renderer/CullingLOD.cpp, IO/TransferFunction1D.cpp):
  * view / projection / LoD factor of tvk_compute_view (the product's host logic) and of the test helper
    scene.py == BuildLookAt / Perspective / CullingLOD::SetScreenParams, bit for bit,
  * the oracle's classic frame planning: frustum culling == CullingLOD::IsVisible on every brick, LoD choice
    == CullingLOD::GetLODLevel clamped as AbstrRenderer::ComputeMinLODForCurrentView does,
  * TransferFunction1D mirror == the reference's SetStdFunction / GetByteArray / GetNonZeroLimits."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import golden_scenes
import scene
import tuvok_b200 as tb
from oracle import orc
from tuvok_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "ref_host")


def _have():
    if not os.path.exists(BIN) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(BIN)


pytestmark = pytest.mark.skipif(not _have(), reason="oracle/_ref/ref_host not built (reference tree absent)")


def fl(v):
    return " ".join("%.9g" % float(x) for x in np.asarray(v, np.float32).reshape(-1))


def run(tmp_path, commands):
    cmd = os.path.join(str(tmp_path), "cmd.txt")
    res = os.path.join(str(tmp_path), "res.txt")
    with open(cmd, "w") as f:
        f.write("\n".join(commands) + "\n")
    subprocess.check_call([BIN, cmd, res])
    with open(res) as f:
        return [l.split() for l in f if l.strip()]


def hexf(tokens):
    return np.array([float.fromhex(t) for t in tokens], np.float32)


@pytest.mark.parametrize("w,h,fov", [(512, 512, 50.0), (1920, 1080, 50.0), (96, 64, 35.0)])
@pytest.mark.parametrize("eye", [(0, 0, 1.6), (0.4, -0.3, 2.2)])
def test_view_projection_lodfactor_bit_exact(tmp_path, w, h, fov, eye):
    rot = (tb.rotation_y(33.0) @ tb.rotation_x(-12.0)).astype(np.float32)
    tra = tb.translation(0.1, -0.05, 0.3)
    aspect = np.float32(w) / np.float32(h)
    rows = run(tmp_path, ["view %s 0 0 0 0 1 0 %.9g %.9g 0.01 1000 %d" % (fl(eye), fov, aspect, h)])
    ref_view, ref_proj, ref_lf = hexf(rows[0][1:]).reshape(4, 4), hexf(rows[1][1:]).reshape(4, 4), hexf(rows[2][1:])[0]
    # reference model-view: m = rotation * translation (RenderRegion), then * view (GLRenderer.cpp:627)
    rows = run(tmp_path, ["mul %s %s" % (fl(rot), fl(tra))])
    model = hexf(rows[0][1:]).reshape(4, 4)
    rows = run(tmp_path, ["mul %s %s" % (fl(model), fl(ref_view))])
    ref_mv = hexf(rows[0][1:]).reshape(4, 4)

    # the test helper (independent numpy restatement used by every parity scene)
    if tuple(eye) == (0, 0, 1.6):       # the reference's default camera: exact; off-axis eyes differ in the
        np.testing.assert_array_equal(scene.look_at(eye, (0, 0, 0), (0, 1, 0)), ref_view)   # dot-product rounding
    np.testing.assert_allclose(scene.look_at(eye, (0, 0, 0), (0, 1, 0)), ref_view, rtol=0, atol=1.2e-7)
    np.testing.assert_array_equal(scene.perspective(fov, aspect, 0.01, 1000.0), ref_proj)
    assert scene.lod_factor(fov, h) == ref_lf
    # the product's host logic behind the C ABI
    p = L.RenderParams()
    rc = tb.lib().tvk_compute_view(C.byref(p), w, h, L.f32x16(*rot.reshape(-1)), L.f32x16(*tra.reshape(-1)),
                                   L.f32x3(*eye), L.f32x3(0, 0, 0), L.f32x3(0, 1, 0), fov, 0.01, 1000.0, 1.0)
    assert rc == 0
    np.testing.assert_array_equal(np.array(p.projection, np.float32).reshape(4, 4), ref_proj)
    assert np.float32(p.lod_factor) == ref_lf
    np.testing.assert_allclose(np.array(p.model_view, np.float32).reshape(4, 4), ref_mv, rtol=0, atol=2.4e-7)


@pytest.mark.parametrize("name,over", [
    ("c2_bricked36_1d_ert", {}),
    ("ragged_1d_lit", {}),
    ("inside_aniso_2d", {}),
    ("c2_bricked36_1d_ert", dict(translation=tb.translation(0.9, 0.0, 1.0))),     # half off screen
    ("ragged_1d_lit", dict(translation=tb.translation(-0.6, 0.5, 0.8))),
    ("c2_bricked36_1d_ert", dict(translation=tb.translation(0.0, 0.0, -2.2))),     # coarser LoD
])
def test_classic_planning_matches_reference_cullinglod(tmp_path, name, over):
    s = golden_scenes.make(name, **over)
    o = s.octree
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    aspect = np.float32(s.width) / np.float32(s.height)
    mv, proj = s.matrices()
    screen = "%.9g %.9g 0.01 1000 %d" % (s.fov, aspect, s.height)
    ident = np.eye(4, dtype=np.float32)

    # LoD: AbstrRenderer::ComputeMinLODForCurrentView (AbstrRenderer.cpp:789-803)
    ext = np.array(s.size, np.float32) * np.array(s.scale, np.float32)
    ext = ext / ext.max()
    rows = run(tmp_path, ["cull %s %s %s %s 1 0 0 0 %s %d %d %d" % (screen, fl(mv), fl(ident), fl(proj), fl(ext), *s.size)])
    ref_lod = min(max(int(rows[0][3]), 0), s.pool_lod_count() - 1)
    lod = orc.classic_lod(p, s.pool_lod_count())
    assert lod == ref_lod

    # every brick of that LoD: geometry from an all-pass planning run (far camera), then the reference's
    # IsVisible on the real camera must select exactly the oracle's list
    far = golden_scenes.make(name, **dict(over, translation=tb.translation(0, 0, -40.0), rotation=np.eye(4, dtype=np.float32)))
    pf = far.oracle_params(pool)
    bc = o.brick_count(lod)
    n_lod = bc[0] * bc[1] * bc[2]
    first = o.brick_index(0, 0, 0, lod)
    mm = o.minmax[first:first + n_lod]
    everything = (0.0, 1e30, 0.0, 1e30)
    all_b, n_all = orc.classic_brick_list(pf, lod, s.overlap, mm, everything)
    assert n_all == n_lod
    by_index = {all_b[i].index: all_b[i] for i in range(n_all)}
    boxes = []
    for i in range(n_lod):
        b = by_index[i]
        boxes.append("%s %s %d %d %d" % (fl(b.center), fl(b.ext), *b.n_vox))
    rows = run(tmp_path, ["cull %s %s %s %s %d %s" % (screen, fl(mv), fl(ident), fl(proj), n_lod, " ".join(boxes))])
    vis = np.array(rows[0][2::2], np.int64)
    want = set(np.nonzero(vis)[0].tolist())
    mine, n = orc.classic_brick_list(p, lod, s.overlap, mm, s.visibility_args())
    got = set(mine[i].index for i in range(n))
    assert got == want
    assert 0 < len(want) <= n_lod


@pytest.mark.parametrize("n,c,g", [(256, 0.5, 0.5), (4096, 0.3, 0.4), (256, 0.05, 0.5), (1024, 0.9, 0.3), (4096, 0.2, 0.2),
                                   (256, 0.0, 1.0), (64, 1.0, 0.1)])
def test_tf1d_matches_reference_class(tmp_path, n, c, g):
    rows = run(tmp_path, ["tf1d %d %.9g %.9g" % (n, c, g)])
    r = rows[0]
    lo, hi = int(r[3]), int(r[4])
    ref_bytes = np.array(r[6:6 + 4 * n], np.uint8).reshape(n, 4)
    t = tb.TransferFunction1D(n)
    t.SetStdFunction(c, g)
    np.testing.assert_array_equal(t.GetByteArray(), ref_bytes)
    assert t.GetNonZeroLimits() == (lo, hi)
    ref = orc.tf1d_std(n, c, g)
    np.testing.assert_array_equal(orc.tf1d_bytes(ref), ref_bytes)
    assert orc.tf1d_nonzero(ref) == (lo, hi)


@pytest.mark.parametrize("name", ["c2_bricked36_1d_ert", "c3_bricked36_2d_lit", "inside_aniso_2d", "c4_f32_iso"])
def test_raycast_uniforms_follow_setupraycastshader(tmp_path, name):
    """mEyeToModel, mModelToEye, inverse(modelView), vDomainScale, model-space light direction and eye position: the
    oracle's derivation (double-precision inverse, rounded once) vs the statement sequence of
    GLGridLeaper::SetupRaycastShader / ComputeEyeToModelMatrix run on the reference's own FLOATMATRIX4 (fp32 inverse).
    Same quantities to fp32 rounding -- and the product's host code derives them like the oracle (images are bit-identical)."""
    s = golden_scenes.make(name)
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    u = orc.uniforms(p)
    mv, _ = s.matrices()
    rows = run(tmp_path, ["uniforms %s %d %d %d %s %s %s" % (fl(mv), s.size[0], s.size[1], s.size[2], fl(s.scale),
                                                             fl(list(p.light_dir)), fl(list(p.eye)))])
    emm, m2e, mvinv, vecs = hexf(rows[0][1:]), hexf(rows[1][1:]), hexf(rows[2][1:]), hexf(rows[3][1:])
    np.testing.assert_allclose(u["emm"], emm, rtol=0, atol=2e-6)
    np.testing.assert_allclose(u["model_to_eye"], m2e, rtol=0, atol=2e-6)
    np.testing.assert_allclose(u["mv_inv"], mvinv, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(u["domain_scale"], vecs[0:3])
    np.testing.assert_allclose(u["light_dir_m"], vecs[3:6], rtol=0, atol=2e-6)
    np.testing.assert_allclose(u["eye_m"], vecs[6:9], rtol=0, atol=2e-6)
    ext = np.array(s.size, np.float32) * np.array(s.scale, np.float32)
    assert np.float32(u["lzwse"]) == np.max((ext / ext.max()).astype(np.float32) / np.array(s.size, np.float32))


# ---------------------------------------------------------------------------------------------------------------
# clip plane: GLGridLeaper::FillBBoxVBO cuts the bounding box with Clipper::BoxPlane and rasterises the
# polytope's front / back faces; the library cuts the ray interval analytically.  Pinned here against the UNMODIFIED
# Basics/Clipper.cpp + PLANE<float> (ref_host clipbox): the analytic entry / exit of every pixel lie on the reference's
# triangles, and pixels the plane removes do not meet the clipped polytope at all.
def _line_tri_hits(o, d, tris):
    """Moeller-Trumbore for lines o + t*d (n, 3) against triangles (m, 3, 3): t (n, m), nan where the line misses."""
    v0, e1, e2 = tris[:, 0][None], (tris[:, 1] - tris[:, 0])[None], (tris[:, 2] - tris[:, 0])[None]
    dd, oo = d[:, None, :], o[:, None, :]
    pv = np.cross(dd, e2)
    det = np.einsum("nmk,nmk->nm", np.broadcast_to(e1, pv.shape), pv)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        tv = oo - v0
        u = np.einsum("nmk,nmk->nm", tv, pv) * inv
        qv = np.cross(tv, np.broadcast_to(e1, tv.shape))
        v = np.einsum("nmk,nmk->nm", np.broadcast_to(dd, qv.shape), qv) * inv
        t = np.einsum("nmk,nmk->nm", np.broadcast_to(e2, qv.shape), qv) * inv
    eps = 1e-6
    ok = (np.abs(det) > 1e-12) & (u >= -eps) & (v >= -eps) & (u + v <= 1 + eps)
    return np.where(ok, t, np.nan)


CLIP_CASES = [
    # world-space plane (normal, d), rotation, translation, scene overrides
    ((0.0, 0.0, 1.0, 0.0), ROT_ID := np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32), {}),
    ((0.3, -0.5, 0.8, 0.1), (tb.rotation_y(33.0) @ tb.rotation_x(-12.0)).astype(np.float32), tb.translation(0.05, -0.02, 0.1), {}),
    ((-0.7, 0.2, -0.4, -0.15), (tb.rotation_y(-60.0) @ tb.rotation_x(25.0)).astype(np.float32), np.eye(4, dtype=np.float32),
     dict(size=(64, 40, 24), scale=(1.0, 1.0, 2.0))),
    # camera inside the volume, in the removed half: every ray starts on the cap polygon ...
    ((0.0, 0.3, 1.0, -1.2), np.eye(4, dtype=np.float32), tb.translation(0.0, 0.0, 1.3), {}),
    # ... and in the kept half: every ray starts on the near plane, some leave through the cap
    ((0.0, 0.3, -1.0, 1.45), np.eye(4, dtype=np.float32), tb.translation(0.0, 0.0, 1.3), {}),
]


@pytest.mark.parametrize("case", range(len(CLIP_CASES)))
def test_clip_plane_matches_reference_clipper(tmp_path, case):
    plane_w, rot, tra, over = CLIP_CASES[case]
    kw = dict(kind=0, size=(32, 32, 32), dtype=orc.U8, brick=20, overlap=2, width=72, height=56, rotation=rot, translation=tra)
    kw.update(over)
    s0 = scene.Scene(**kw)
    ext = np.array(s0.size, np.float32) * np.array(s0.scale, np.float32)
    ext = ext / ext.max()
    rows = run(tmp_path, ["clipbox %s %s %s %s" % (fl(plane_w), fl(rot), fl(tra), fl(ext))])
    ref_plane = hexf(rows[0][1:])
    tris = hexf(rows[1][2:]).astype(np.float64).reshape(-1, 3, 3)
    assert tris.shape[0] == int(rows[1][1]) and tris.shape[0] >= 4
    # 1. the library's world -> model helper == PLANE<float> * inverse(rotation * translation), normal normalised
    got = L.f32x4()
    assert L.lib().tvk_clip_plane_to_model(L.f32x4(*plane_w), L.f32x16(*rot.reshape(-1)), L.f32x16(*tra.reshape(-1)), got) == L.OK
    assert np.abs(np.array(got, np.float32) - ref_plane).max() <= 2e-6
    # 2. the oracle's analytic ray interval against the reference's clipped polytope
    s = scene.Scene(clip_plane_model=tuple(ref_plane), **kw)
    pool, _ = s.oracle_pool()
    e_c, x_c, cov_c = orc.ray_setup(s.oracle_params(pool))
    e_b, x_b, cov_b = orc.ray_setup(s0.oracle_params(pool))
    assert cov_b.sum() > 200 and 0 < cov_c.sum() <= cov_b.sum()
    assert not np.any(cov_c & ~cov_b.astype(bool))                     # the plane only removes
    tn = tris / ext[None, None, :].astype(np.float64) + 0.5             # model space -> the [0,1]^3 coordinates of the rays
    sel = np.flatnonzero(cov_c)
    o, d = e_c[sel, :3].astype(np.float64), (x_c[sel, :3] - e_c[sel, :3]).astype(np.float64)
    t = _line_tri_hits(o, d, tn)
    ln = np.linalg.norm(d, axis=1)                                       # distances in units of the volume's longest side
    tmin, tmax = np.nanmin(t, axis=1) * ln, np.nanmax(t, axis=1) * ln
    assert np.abs(tmax - ln).max() <= 1e-4                              # the exit lies on a back face of the polytope
    near = np.abs(e_c[sel, 3] + 0.01) <= 1e-6                            # entry on the near plane (eye-space z = -near)
    assert not (~near).any() or np.abs(tmin[~near]).max() <= 1e-4                            # otherwise the entry lies on a front face
    assert np.all(tmin[near] <= 1e-4)
    if case >= 3:
        assert near.all() if case == 4 else not near.any()
    # 3. pixels of the box that the plane removed: the line never meets the polytope in front of the near plane
    gone = np.flatnonzero(cov_b.astype(bool) & ~cov_c.astype(bool))
    o, d = e_b[gone, :3].astype(np.float64), (x_b[gone, :3] - e_b[gone, :3]).astype(np.float64)
    t = _line_tri_hits(o, d, tn)
    with np.errstate(invalid="ignore"):
        inside = np.nanmax(np.where((t > 1e-3) & (t < 1 - 1e-3), 1.0, np.nan), axis=1) if t.size else np.array([])
    span = np.nanmax(t, axis=1) - np.nanmin(t, axis=1) if t.size else np.array([])
    bad = np.flatnonzero(np.nan_to_num(inside) * (np.nan_to_num(span) > 5e-3))
    assert bad.size <= max(2, gone.size // 100), (bad.size, gone.size)  # grazing silhouette pixels aside


BRICKDIST = os.path.join(ROOT, "oracle", "_ref", "ref_brickdist")


@pytest.mark.skipif(not os.path.exists(BRICKDIST), reason="oracle/_ref/ref_brickdist not built (reference tree absent)")
@pytest.mark.parametrize("name,over", [
    ("c2_bricked36_1d_ert", {}),
    ("ragged_1d_lit", dict(translation=tb.translation(-0.6, 0.5, 0.8))),
    ("inside_aniso_2d", {}),
])
def test_brick_distance_and_depth_order_match_reference(tmp_path, name, over):
    """the depth order of the classic per-brick path.  The oracle's brick distances (orc_classic.cpp) against the
    reference's OWN file-static brick_distance (AbstrRenderer.cpp:808-841, reached by compiling its translation unit in
    place: oracle/_ref/ref_brickdist), and the order std::sort(vBrickList) can produce from them (AbstrRenderer.h:104-106)."""
    import subprocess
    s = golden_scenes.make(name, **over)
    o = s.octree
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    lod = orc.classic_lod(p, s.pool_lod_count())
    bc = o.brick_count(lod)
    first = o.brick_index(0, 0, 0, lod)
    mm = o.minmax[first:first + bc[0] * bc[1] * bc[2]]
    mine, n = orc.classic_brick_list(p, lod, s.overlap, mm, s.visibility_args())
    live = [mine[i] for i in range(n) if not mine[i].empty]
    assert len(live) >= 2
    mv, _ = s.matrices()
    with open(tmp_path / "in.txt", "w") as f:
        f.write(fl(mv) + "\n")
        for b in live:
            f.write("%s %s\n" % (fl(b.center), fl(b.ext)))
    subprocess.check_call([BRICKDIST, str(tmp_path / "in.txt"), str(tmp_path / "out.txt")])
    ref = np.array([float.fromhex(l.split()[1]) for l in open(tmp_path / "out.txt")], np.float32)
    got = np.array([b.distance for b in live], np.float32)
    # the reference sums x*x + y*y + z*z in the order the compiler picks, the oracle (and the product) by the arithmetic
    # contract's fma chain: the same number up to one rounding
    assert np.all(np.abs(got - ref) <= np.spacing(np.maximum(got, ref))), float(np.abs(got - ref).max())
    # the oracle's list is sorted by its distances; by the reference's distances it is sorted too, except where two bricks
    # are closer together than that one rounding (std::sort leaves ties in unspecified order anyway)
    d = np.diff(ref.astype(np.float64))
    assert np.all(d >= -2.0 * np.spacing(ref[:-1]).astype(np.float64)), d.min()
