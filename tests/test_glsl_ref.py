"""Image-side parity pinned to the reference's OWN shader code: the GridLeaper fragment shader
(Shaders/GLGridLeaper-blend.glsl + Method-{1D,1D-L,2D,2D-L} + GradientTools + lighting + Compositing), the GLSL generated
by the unmodified GLVolumePool (GetBrick page-table walk, ComputeLOD, TransformToPoolSpace, samplePool) and by
GLHashTable (miss reports) are compiled as C++ against oracle/glsl/glsl_emu.h (a minimal GLSL emulation with IEEE fp32
semantics) and EXECUTED per fragment (tests/glsl_ref.py) -- on the same page table, pool atlas, transfer function, ray
entry points and uniforms as the oracle restatement (oracle/orc_render.c), which the CUDA kernel is compared with bit
for bit on the GPU.

What must agree: every brick decision and miss report (integer, exact), the sample counts, and the float images to
rounding (the shader text compiled by g++ evaluates `a*b+c` unfused where the contract fuses the compositing update and
takes gradient taps through texture coordinates instead of texel offsets): |delta| of the RGBA32F image <= 5e-5 on >= 99 %
of the pixels (a few early-terminated rays stop one sample apart: <= 4e-3), RGBA8 images identical or 1/255 apart -- far inside the 2/255 / 45 dB budget of BASELINE.json."""
import numpy as np
import pytest

import glsl_ref
import golden_scenes
import tuvok_b200 as tb
from oracle import orc
from scene import Scene, image_diff
from tuvok_b200 import synth

pytestmark = pytest.mark.skipif(not glsl_ref.available(), reason="reference shaders / oracle/_ref tools absent")


def execute_reference_glsl(tmp_path, s, state, ray_start, start_color, meta):
    p = state["params"]
    o = s.octree
    pool = state["pool"]
    pool_glsl, hash_glsl = glsl_ref.generated_glsl(tmp_path, o, s.size, s.brick[0], s.overlap, s.dtype, pool.pool_size,
                                                   s.strategy, o.brick_count(0), p.hash_size, p.rehash_count)
    exe = glsl_ref.build(tmp_path, s.mode, s.lighting, pool_glsl, hash_glsl)
    u = orc.uniforms(p)
    return glsl_ref.run(exe, tmp_path, p, u["emm"], orc.ray_exit_eye(p), ray_start, start_color, state["covered"], meta,
                        pool.meta_dim, state["atlas"], state["tf"], u["norm"], u, u["domain_scale"])


SCENES = ["c1_single_brick_1d", "c2_bricked36_1d_ert", "c3_bricked36_2d_lit", "ragged_1d_lit", "inside_aniso_2d"]


@pytest.mark.parametrize("name", SCENES)
def test_converged_frame_matches_executed_reference_shader(tmp_path, name):
    s = golden_scenes.make(name)
    st = s.oracle_render()                                    # paging loop until converged
    p = st["params"]
    zeros = np.zeros_like(st["entry"])
    hash_o = np.zeros(p.hash_size, np.uint32)
    outs, rs = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], hash_o, 1)
    g0, g1, g2, hash_g = execute_reference_glsl(tmp_path, s, st, st["entry"], zeros, st["meta"])
    assert not hash_g.any() and not hash_o.any()              # converged: nothing is reported missing
    a, b = outs[0].reshape(-1, 4), g0
    d = np.abs(a - b).max(axis=1)
    # rounding-level agreement everywhere, except where early ray termination (alpha > 0.99) trips one sample apart
    # because the shader text's compositing update is unfused: such a pixel differs by at most that sample's
    # contribution (< 0.01 of a colour), and there are only a handful of them
    assert float(d.max()) <= 4e-3 and float((d > 5e-5).mean()) <= 0.01, (float(d.max()), float((d > 5e-5).mean()))
    print("executed reference GLSL vs oracle: max |d| %.3g, pixels > 5e-5: %d of %d" % (d.max(), int((d > 5e-5).sum()), len(d)))
    mx, psnr = image_diff(orc.rgba8(a.reshape(s.height, s.width, 4)), orc.rgba8(b.reshape(s.height, s.width, 4)))
    assert mx <= 1 and psnr >= 60.0
    # resume buffers of a converged frame: every covered ray finished (w = 1000), resume colour = final colour
    cov = st["covered"].reshape(-1).astype(bool)
    assert (g2[cov, 3] == 1000.0).all() and np.array_equal(outs[2].reshape(-1, 4)[:, 3], g2[:, 3])
    assert float(np.abs(outs[1].reshape(-1, 4) - g1).max()) <= 4e-3
    assert (b[~cov] == 0).all()


@pytest.mark.parametrize("name", ["c2_bricked36_1d_ert", "c3_bricked36_2d_lit", "inside_aniso_2d"])
def test_first_pass_miss_reports_and_resume_state(tmp_path, name):
    """Only the coarsest brick is resident: every ray falls back to it and reports what it misses.  The miss table
    (integer, sequential execution in pixel order on both sides) must be IDENTICAL, and so must the resume positions'
    bookkeeping (which rays stopped being optimal and where)."""
    s = golden_scenes.make(name, hash_size=509)
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    o = s.octree
    ps = pool.pool_size
    atlas = np.zeros((ps[2], ps[1], ps[0]), orc.NP_DTYPE[s.dtype])
    cap = pool.capacity
    last = cap[0] * cap[1] * cap[2] - 1
    b = o.brick(0, 0, 0, pool.lod_count - 1)
    sx, sy, sz = last % cap[0], (last // cap[0]) % cap[1], last // (cap[0] * cap[1])
    atlas[sz * s.brick[2]:sz * s.brick[2] + b.shape[0], sy * s.brick[1]:sy * s.brick[1] + b.shape[1],
          sx * s.brick[0]:sx * s.brick[0] + b.shape[2]] = b
    entry, exit_, cov = orc.ray_setup(p)
    zeros = np.zeros_like(entry)
    hash_o = np.zeros(p.hash_size, np.uint32)
    tf = s.tf_bytes()
    outs, rs = orc.raycast(p, atlas, pool.meta, tf, entry, zeros, exit_, cov, hash_o, 1)
    state = dict(params=p, pool=pool, atlas=atlas, tf=tf, covered=cov)
    g0, g1, g2, hash_g = execute_reference_glsl(tmp_path, s, state, entry, zeros, pool.meta)
    assert hash_o.any()
    assert np.array_equal(hash_g, hash_o)                                     # miss reports: bit-exact
    assert np.array_equal(orc.hash_decode(hash_g, o.brick_count(0)), orc.hash_decode(hash_o, o.brick_count(0)))
    o2, o1, o0 = outs[2].reshape(-1, 4), outs[1].reshape(-1, 4), outs[0].reshape(-1, 4)
    assert np.array_equal(o2[:, 3] == 1000.0, g2[:, 3] == 1000.0)             # the same rays ended optimally
    assert float(np.abs(o2 - g2).max()) <= 1e-5                               # resume positions
    assert float(np.abs(o0 - g0).max()) <= 2e-5 and float(np.abs(o1 - g1).max()) <= 2e-5


def test_sample_rate_and_ragged_bricks(tmp_path):
    """sampleRateModifier != 1 (opacity correction through pow) on ragged u8 bricks with lighting."""
    s = golden_scenes.make("ragged_1d_lit", sample_rate=1.7)
    st = s.oracle_render()
    p = st["params"]
    zeros = np.zeros_like(st["entry"])
    outs, _ = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], None, 1)
    g0, _, _, _ = execute_reference_glsl(tmp_path, s, st, st["entry"], zeros, st["meta"])
    assert float(np.abs(outs[0].reshape(-1, 4) - g0).max()) <= 5e-5


def test_isosurface_shader_and_deferred_compose(tmp_path):
    """GLGridLeaper-iso.glsl (first hit, 5-step refinement, eye-space hit position, normal through mModelViewIT, resume
    encoding) and Compose-FS.glsl (deferred lighting) executed vs the oracle's restatement, f32 volume."""
    s = golden_scenes.make("c4_f32_iso")
    st = s.oracle_render()
    p = st["params"]
    o = s.octree
    pool = st["pool"]
    zeros = np.zeros_like(st["entry"])
    outs, _ = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], None, 1)
    pool_glsl, hash_glsl = glsl_ref.generated_glsl(tmp_path, o, s.size, s.brick[0], s.overlap, s.dtype, pool.pool_size,
                                                   s.strategy, o.brick_count(0), p.hash_size, p.rehash_count)
    exe = glsl_ref.build_iso(tmp_path, pool_glsl, hash_glsl)
    u = orc.uniforms(p)
    g, hash_g = glsl_ref.run_iso(exe, tmp_path, p, u, orc.ray_exit_eye(p), st["entry"], zeros, st["covered"], st["meta"],
                                 pool.meta_dim, st["atlas"])
    assert not hash_g.any()
    hit_o, nrm_o = outs[0].reshape(-1, 4), outs[1].reshape(-1, 4)
    assert np.array_equal(hit_o[:, 3] != 0, g[0][:, 3] != 0)                  # the same rays hit the surface
    assert (hit_o[:, 3] != 0).sum() > 300
    assert float(np.abs(hit_o - g[0]).max()) <= 2e-5                          # eye-space hit positions
    assert float(np.abs(nrm_o - g[1]).max()) <= 2e-4                          # normals (normalised gradients)
    assert np.array_equal(outs[2].reshape(-1, 4)[:, 3], g[2][:, 3])           # resume encoding 1000 / 499 + alpha
    # deferred compose: the reference shader on the REFERENCE shader's buffers vs the oracle on the oracle's
    cexe = glsl_ref.build_compose(tmp_path)
    amb = [p.ambient[i] * p.ambient[3] for i in range(3)]
    dif = [p.diffuse[i] * p.diffuse[3] * p.iso_color[i] for i in range(3)]
    spe = [p.specular[i] * p.specular[3] for i in range(3)]
    img_g = glsl_ref.run_compose(cexe, tmp_path, s.width, s.height, amb, dif, spe, list(p.light_dir), g[0], g[1])
    img_o = orc.iso_compose(p, outs[0], outs[1]).reshape(-1, 4)
    assert float(np.abs(img_o - img_g).max()) <= 2e-4
    mx, psnr = image_diff(orc.rgba8(img_o.reshape(s.height, s.width, 4)), orc.rgba8(img_g.reshape(s.height, s.width, 4)))
    assert mx <= 1 and psnr >= 60.0


CLASSIC = [("c2_bricked36_1d_ert", {}), ("ragged_1d_lit", {}), ("inside_aniso_2d", {}),
           ("ragged_1d_lit", dict(mode=orc.RM_2DTRANS, lighting=True)),
           ("c2_bricked36_1d_ert", dict(translation=tb.translation(0.0, 0.0, -2.2)))]      # coarser LoD: fStepScale = 2


@pytest.mark.parametrize("name,over", CLASSIC)
def test_classic_per_brick_shaders_executed(tmp_path, name, over):
    """the classic GLRaycaster fragment shaders (GLRaycaster-{1D,1D-light,2D,2D-light}-FS.glsl with
    VRender1D*.glsl, Volume3D.glsl, lighting.glsl, Compositing.glsl) executed per brick in the oracle's brick order,
    with the per-brick pass setup (RGBA16F entry FBO, back-face fragments, eye->texture matrix in gl_TextureMatrix[0],
    gl_NormalMatrix, GL under-blending) supplied by the driver, vs orc_classic_render."""
    s = golden_scenes.make(name, **over)
    ref = s.oracle_classic()
    o = s.octree
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    lod = ref["lod"]
    bc = o.brick_count(lod)
    first = o.brick_index(0, 0, 0, lod)
    mm = o.minmax[first:first + bc[0] * bc[1] * bc[2]]
    bricks, n = orc.classic_brick_list(p, lod, s.overlap, mm, s.visibility_args())
    data = [None if bricks[i].empty else o.brick(*bricks[i].coord, lod) for i in range(n)]
    u = orc.uniforms(p)
    light = dict(ambient=u["ambient"], diffuse=u["diffuse"], specular=u["specular"], dir=list(p.light_dir))
    exe = glsl_ref.build_classic(tmp_path, s.mode, s.lighting)
    img = glsl_ref.run_classic(exe, tmp_path, p, u["inv_proj"], u["mv_inv"], orc.classic_step_scale(p, lod), u["norm"],
                               u["domain_scale"], light, bricks, n, data, s.tf_bytes())
    a = ref["image"].reshape(-1, 4)
    d = np.abs(a - img).max(axis=1)
    print("classic shaders executed vs oracle: max |d| %.3g, pixels > 5e-5: %d of %d" % (d.max(), int((d > 5e-5).sum()), len(d)))
    assert float(d.max()) <= 8e-3 and float((d > 5e-5).mean()) <= 0.01, (float(d.max()), float((d > 5e-5).mean()))
    mx, psnr = image_diff(orc.rgba8(a.reshape(s.height, s.width, 4)), orc.rgba8(img.reshape(s.height, s.width, 4)))
    assert mx <= 2 and psnr >= 55.0


def test_openmp_baseline_binary_equals_sequential_execution(tmp_path):
    """bench.py's cpu_baseline of kind "reference" (glsl_ref.build_baseline: the same shader text, per-fragment globals
    thread_local, fragments distributed with OpenMP) renders bit for bit what the sequential executed shader renders."""
    import os
    import subprocess
    s = golden_scenes.make("c3_bricked36_2d_lit")
    st = s.oracle_render()
    p, pool = st["params"], st["pool"]
    zeros = np.zeros_like(st["entry"])
    g0, g1, g2, _ = execute_reference_glsl(tmp_path, s, st, st["entry"], zeros, st["meta"])
    pool_glsl, hash_glsl = glsl_ref.generated_glsl(tmp_path, s.octree, s.size, s.brick[0], s.overlap, s.dtype, pool.pool_size,
                                                   s.strategy, s.octree.brick_count(0), p.hash_size, p.rehash_count)
    exe = glsl_ref.build_baseline(tmp_path, os.path.join(str(tmp_path), "baseline"), s.mode, s.lighting, pool_glsl, hash_glsl)
    u = orc.uniforms(p)
    fin, fout = os.path.join(str(tmp_path), "b_scene.bin"), os.path.join(str(tmp_path), "b_out.bin")
    glsl_ref.scene_file(fin, p, u, orc.ray_exit_eye(p), st["entry"], zeros, st["covered"], st["meta"], pool.meta_dim,
                        st["atlas"], st["tf"])
    out = subprocess.run([exe, fin, fout, "2"], capture_output=True, text=True, check=True,
                         env=dict(os.environ, OMP_NUM_THREADS="4")).stdout.split()
    assert float(out[0]) > 0.0
    npx = p.width * p.height
    img = np.fromfile(fout, np.float32, count=npx * 12).reshape(3, npx, 4)
    assert np.array_equal(img[0], g0) and np.array_equal(img[1], g1) and np.array_equal(img[2], g2)


CLASSIC_ISO = [("c4_f32_iso", {}), ("c2_bricked36_1d_ert", dict(mode=orc.RM_ISOSURFACE)),
               ("ragged_1d_lit", dict(mode=orc.RM_ISOSURFACE, isovalue=90.0)), ("inside_aniso_2d", dict(mode=orc.RM_ISOSURFACE))]


@pytest.mark.parametrize("name,over", CLASSIC_ISO)
def test_classic_isosurface_shaders_executed(tmp_path, name, over):
    """GLRaycaster-ISO-FS.glsl + RefineIsosurface.glsl + Volume3D.glsl executed per brick in the oracle's
    brick order with the RM_ISOSURFACE pass setup of GLRaycaster::Render3DInLoop (two float targets, gl_FragDepth under
    DF_LESS), vs orc_classic_iso_render: the same pixels hit, the same brick wins the depth test, positions and normals
    to rounding (a bisection step of RefineIsosurface may flip on the noise volume: <= 1/32 of a sample step)."""
    s = golden_scenes.make(name, **over)
    r = s.oracle_classic()
    p = r["params"]
    u = orc.uniforms(p)
    pr = np.array(list(p.projection), np.float64)
    zn, zf = pr[14] / (pr[10] - 1.0), pr[14] / (pr[10] + 1.0)
    pp = (np.float32(zf / (zf - zn)), np.float32(zf * zn / (zn - zf)))
    exe = glsl_ref.build_classic_iso(tmp_path)
    hp, hn = glsl_ref.run_classic_iso(exe, tmp_path, p, u["inv_proj"], u["mv_inv"], u["norm"], u["domain_scale"], pp, r["bricks"],
                                      r["n"], r["data"])
    a, b = r["hit_pos"], r["hit_normal"]
    hit_o, hit_g = a[:, 3] != 0, hp[:, 3] != 0
    assert np.array_equal(hit_o, hit_g)
    if hit_o.any():
        assert np.array_equal(b[hit_o, 3], hn[hit_o, 3])                       # iTileID of the winning brick
        assert float(np.abs(a[hit_o] - hp[hit_o]).max()) <= 2e-3
        assert float(np.abs(b[hit_o, :3] - hn[hit_o, :3]).max()) <= 2e-2
        assert float((np.abs(a[hit_o] - hp[hit_o]).max(axis=1) > 1e-5).mean()) <= 0.02
    img_o = orc.iso_compose(p, a, b)
    img_g = orc.iso_compose(p, hp, hn)
    mx, psnr = image_diff(orc.rgba8(img_o.reshape(s.height, s.width, 4)), orc.rgba8(img_g.reshape(s.height, s.width, 4)))
    assert mx <= 2 and psnr >= 45.0


CLASSIC_CV = [("c4_f32_iso", {}, 0.8), ("c2_bricked36_1d_ert", dict(mode=orc.RM_ISOSURFACE), 45000.0),
              ("ragged_1d_lit", dict(mode=orc.RM_ISOSURFACE, isovalue=90.0), 140.0)]


@pytest.mark.parametrize("name,over,cv_iso", CLASSIC_CV)
def test_clearview_shaders_executed(tmp_path, name, over, cv_iso):
    """GLRaycaster-ISO-FS.glsl and, right after it per brick, GLRaycaster-ISO-CV-FS.glsl (reading the first
    pass's targets as texLastHit / texLastHitPos, depth of the ray exit under DF_LESS), then Compose-CV-FS.glsl, vs
    orc_classic_cv_render + orc_cv_compose: the same pixels hit in both passes, the same bricks win, values to rounding."""
    s = golden_scenes.make(name, **over)
    s.clearview = dict(isovalue=cv_iso, size=2.5)
    r = s.oracle_classic()
    p = r["params"]
    u = orc.uniforms(p)
    pr = np.array(list(p.projection), np.float64)
    zn, zf = pr[14] / (pr[10] - 1.0), pr[14] / (pr[10] + 1.0)
    pp = (np.float32(zf / (zf - zn)), np.float32(zf * zn / (zn - zf)))
    d = np.array(u["diffuse"], np.float32)
    light = dict(ambient=u["ambient"], diffuse=d * np.array(list(p.iso_color), np.float32), specular=u["specular"],
                 dir=list(p.light_dir))
    d2 = d * np.array([1.0, 0.0, 0.0], np.float32)
    exe = glsl_ref.build_classic_cv(tmp_path)
    out = glsl_ref.run_classic_cv(exe, tmp_path, p, u["inv_proj"], u["mv_inv"], u["norm"], u["domain_scale"], pp, light,
                                  r["cv_isoval"], d2, r["cv_param"], r["pick"], r["bricks"], r["n"], r["data"])
    for k, (pos, nrm) in enumerate((("hit_pos", "hit_normal"), ("cv_pos", "cv_normal"))):
        a, b, an, bn = r[pos], out[2 * k], r[nrm], out[2 * k + 1]
        hit = a[:, 3] != 0
        assert np.array_equal(hit, b[:, 3] != 0)
        if hit.any():
            assert np.array_equal(an[hit, 3], bn[hit, 3])                      # iTileID of the brick that won the depth test
            assert float(np.abs(a[hit] - b[hit]).max()) <= 2e-3 and float(np.abs(an[hit, :3] - bn[hit, :3]).max()) <= 2e-2
    assert r["cv_pos"][:, 3].any()
    img = r["image"].reshape(-1, 4)
    assert float(np.abs(img - out[4]).max()) <= 5e-3 and float((np.abs(img - out[4]).max(axis=1) > 1e-4).mean()) <= 0.01
    mx, psnr = image_diff(orc.rgba8(r["image"]), orc.rgba8(out[4].reshape(s.height, s.width, 4)))
    assert mx <= 2 and psnr >= 45.0
