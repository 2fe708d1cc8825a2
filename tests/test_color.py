"""Colour (RGBA, 4 x 8 bit) volumes on the GridLeaper path: GLGridLeaper-Method-{1D,1D-L,2D,2D-L,iso}-
color.glsl (the volume's own colour, the transfer function maps alpha only; ComputeGradientAlpha; the 1D-L / iso normals
from the .r channel as the shader text has it) and Compose-Color-FS.glsl (hit colour packed into the two alpha channels).

  * CPU: the oracle's restatement (orc_render.c, dtype ORC_RGBA8) against the reference's OWN shader text executed through
    oracle/glsl/glsl_emu.h -- the colour method files linked with the unmodified blend / iso main shaders and the GLSL the
    unmodified GLVolumePool / GLHashTable generate,
  * (-m gpu) the CUDA colour kernel (csrc/k_color.cu) against the oracle: float images, resume buffers, miss lists and
    page tables bit-identical in all five modes."""
import numpy as np
import pytest

import glsl_ref
import tuvok_b200 as tb
from oracle import orc
from scene import ColorScene, image_diff
from tuvok_b200 import synth

ROT = (tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32)
MODES = [(orc.RM_1DTRANS, False), (orc.RM_1DTRANS, True), (orc.RM_2DTRANS, False), (orc.RM_2DTRANS, True)]


def scene(mode, lighting, **kw):
    args = dict(kind=synth.V_SPH, size=(48, 40, 36), brick=16, overlap=2, mode=mode, lighting=lighting, width=64, height=48,
                rotation=ROT, isovalue=90, tf_center=0.3, tf_inv_gradient=0.35)
    args.update(kw)
    return ColorScene(**args)


def test_colour_octree_is_four_scalar_conversions():
    s = scene(orc.RM_1DTRANS, False)
    o = s.octree
    a = orc.Octree(np.ascontiguousarray(s.volume[..., 3]), s.brick, s.overlap)
    assert np.array_equal(o.minmax, a.minmax)                      # visibility sees the alpha channel (uvfDataset.cpp:1144)
    b = o.brick(1, 0, 1, 0)
    assert b.shape[3] == 4 and np.array_equal(b[..., 3], a.brick(1, 0, 1, 0))
    assert np.array_equal(b[2:-2, 2:-2, 2:-2, 0], s.volume[12:24, 0:12, 12:24, 0])     # inner voxels of brick (1, 0, 1)


def _golden_colour_file():
    import importlib.util
    import os
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_octree_golden", os.path.join(golden, "make_octree_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return os.path.join(golden, "octree_rgba8_zlib.bin"), m.CASES["octree_rgba8_zlib"], m.volume("octree_rgba8_zlib")


def test_colour_octree_file_written_by_the_reference_converter_reads_back():
    """tests/golden/octree_rgba8_zlib.bin: an ExtendedOctree file with four 8-bit components written by the unmodified
    converter (zlib bricks).  The product's file reader (host side, no device) delivers the interleaved bricks of the oracle."""
    from tuvok_b200 import octree_file
    path, (kind, size, dt, _, brick, ov, _, _), vol = _golden_colour_file()
    info = octree_file.probe(path)
    assert info.dtype == tb.RGBA8 and tuple(info.domain_size) == tuple(size) and info.overlap == ov
    o = orc.ColorOctree(vol, brick, ov)
    assert info.brick_count == o.total_bricks and info.lod_count == o.lod_count
    n = 0
    for key in o.iter_bricks():
        got = octree_file.read_brick(path, *key, info=info)
        assert got.shape[3] == 4 and np.array_equal(got, o.brick(*key)), key
        n += 1
    assert n == o.total_bricks


def test_colour_uvf_container_walk_takes_the_alpha_minmax():
    """tests/golden/volume_rgba8_zlib.uvf: a complete UVF written by the reference's own UVF / TOCBlock / MaxMinDataBlock
    classes for four-component data.  The product's container walk pairs the TOC block with the MaxMin block and takes
    component 3 of every entry, as UVFDataset::MaxMinForKey does (uvfDataset.cpp:1188)."""
    import os
    from tuvok_b200 import octree_file
    path, (kind, size, dt, _, brick, ov, _, _), vol = _golden_colour_file()
    uvf = os.path.join(os.path.dirname(path), "volume_rgba8_zlib.uvf")
    u = octree_file.uvf_probe(uvf)
    o = orc.ColorOctree(vol, brick, ov)
    assert u["n_timesteps"] == 1 and u["maxmin"].shape == (o.total_bricks, 4)
    assert np.array_equal(u["maxmin"][:, :2], o.minmax[:, :2])
    info = octree_file.probe(uvf, offset=u["toc_payload_offset"], uvf_file_version=u["file_version"])
    assert info.dtype == tb.RGBA8 and info.brick_count == o.total_bricks
    assert np.array_equal(octree_file.read_brick(uvf, 1, 2, 1, 0, info=info, offset=u["toc_payload_offset"]), o.brick(1, 2, 1, 0))


@pytest.mark.skipif(not __import__("os").path.exists(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))), "oracle", "_ref", "ref_dataset")),
                    reason="oracle/_ref/ref_dataset not built (reference tree absent)")
def test_reference_uvfdataset_sees_the_same_colour_dataset(tmp_path):
    """The unmodified UVFDataset on the colour golden: four components, and MaxMinForKey = the oracle's alpha statistics."""
    import os
    import subprocess
    path, (kind, size, dt, _, brick, ov, _, _), vol = _golden_colour_file()
    uvf = os.path.join(os.path.dirname(path), "volume_rgba8_zlib.uvf")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "ds.txt"
    subprocess.check_call([os.path.join(root, "oracle", "_ref", "ref_dataset"), uvf, str(out), "256"], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    o = orc.ColorOctree(vol, brick, ov)
    rows = [l.split() for l in open(out)]
    head = rows[0]
    assert int(head[head.index("comps") + 1]) == 4 and int(head[head.index("bits") + 1]) == 8
    n = 0
    for r in rows:
        if r[0] != "brick":
            continue
        lod, idx = int(r[1]), int(r[2])
        bc = o.brick_count(lod)
        x, y, z = idx % bc[0], (idx // bc[0]) % bc[1], idx // (bc[0] * bc[1])
        i = o.brick_index(x, y, z, lod)
        k = r.index("mm")
        assert (float.fromhex(r[k + 1]), float.fromhex(r[k + 2])) == (o.minmax[i, 0], o.minmax[i, 1]), (lod, idx)
        n += 1
    assert n == o.total_bricks


needs_glsl = pytest.mark.skipif(not glsl_ref.available(), reason="reference shaders / oracle/_ref tools absent")


def _generated(tmp_path, s, pool, p):
    o = s.octree
    # the page-table walk / hash GLSL depends on the pool geometry only; the alpha channel's octree stands in for the data
    return glsl_ref.generated_glsl(tmp_path, o.ch[3], s.size, s.brick[0], s.overlap, orc.U8, pool.pool_size, s.strategy,
                                   o.brick_count(0), p.hash_size, p.rehash_count)


@needs_glsl
@pytest.mark.parametrize("mode,lighting", MODES)
def test_colour_methods_match_executed_reference_shaders(tmp_path, mode, lighting):
    s = scene(mode, lighting)
    st = s.oracle_render()
    p, pool = st["params"], st["pool"]
    zeros = np.zeros_like(st["entry"])
    hash_o = np.zeros(p.hash_size, np.uint32)
    outs, _ = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], hash_o, 1)
    pool_glsl, hash_glsl = _generated(tmp_path, s, pool, p)
    exe = glsl_ref.build(tmp_path, s.mode, s.lighting, pool_glsl, hash_glsl, color=True)
    u = orc.uniforms(p)
    g0, g1, g2, hash_g = glsl_ref.run(exe, tmp_path, p, u["emm"], orc.ray_exit_eye(p), st["entry"], zeros, st["covered"],
                                      st["meta"], pool.meta_dim, st["atlas"], st["tf"], u["norm"], u, u["domain_scale"])
    # the pool of the converged frame holds every brick its RESUMED rays asked for; a fresh pass from the entry points (this
    # one) can end a ray a sample later (resume arithmetic / unfused compositing update) and ask for one brick more
    assert np.count_nonzero(hash_o) <= 2 and np.count_nonzero(hash_g) <= 2
    a, b = outs[0].reshape(-1, 4), g0
    assert a[:, 3].max() > 0.5 and (a[:, :3].max(axis=0) > 0.05).all()          # something coloured is on screen
    d = np.abs(a - b).max(axis=1)
    # rounding-level agreement; a pixel whose early termination trips one sample apart differs by that sample (as for scalar
    # volumes, tests/test_glsl_ref.py)
    assert float(d.max()) <= 4e-3 and float((d > 5e-5).mean()) <= 0.01, (float(d.max()), float((d > 5e-5).mean()))
    mx, psnr = image_diff(orc.rgba8(a.reshape(s.height, s.width, 4)), orc.rgba8(b.reshape(s.height, s.width, 4)))
    assert mx <= 1 and psnr >= 60.0
    assert int((outs[2].reshape(-1, 4)[:, 3] != g2[:, 3]).sum()) <= 2          # resume depth: all but those one or two rays


@needs_glsl
def test_colour_first_pass_reports_the_same_missing_bricks(tmp_path):
    s = scene(orc.RM_2DTRANS, True)
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    atlas = np.zeros((pool.pool_size[2], pool.pool_size[1], pool.pool_size[0]), orc.NP_DTYPE[orc.RGBA8])
    o = s.octree
    b = o.brick(0, 0, 0, pool.lod_count - 1)
    cap = pool.capacity
    z0, y0, x0 = (cap[2] - 1) * s.brick[2], (cap[1] - 1) * s.brick[1], (cap[0] - 1) * s.brick[0]
    atlas[z0:z0 + b.shape[0], y0:y0 + b.shape[1], x0:x0 + b.shape[2]] = b
    entry, exit_, cov = orc.ray_setup(p)
    zeros = np.zeros_like(entry)
    hash_o = np.zeros(p.hash_size, np.uint32)
    outs, _ = orc.raycast(p, atlas, pool.meta, s.tf_bytes(), entry, zeros, exit_, cov, hash_o, 1)
    pool_glsl, hash_glsl = _generated(tmp_path, s, pool, p)
    exe = glsl_ref.build(tmp_path, s.mode, s.lighting, pool_glsl, hash_glsl, color=True)
    u = orc.uniforms(p)
    g0, g1, g2, hash_g = glsl_ref.run(exe, tmp_path, p, u["emm"], orc.ray_exit_eye(p), entry, zeros, cov, pool.meta,
                                      pool.meta_dim, atlas, s.tf_bytes(), u["norm"], u, u["domain_scale"])
    assert hash_o.any() and np.array_equal(np.sort(hash_o), np.sort(hash_g))      # the same bricks are requested
    assert float(np.abs(outs[0].reshape(-1, 4) - g0).max()) <= 4e-3
    assert np.array_equal(outs[2].reshape(-1, 4)[:, 3] == 1000.0, g2[:, 3] == 1000.0)


@needs_glsl
def test_colour_isosurface_and_compose_match_executed_reference_shaders(tmp_path):
    s = scene(orc.RM_ISOSURFACE, False)
    st = s.oracle_render()
    p, pool = st["params"], st["pool"]
    zeros = np.zeros_like(st["entry"])
    outs, _ = orc.raycast(p, st["atlas"], st["meta"], st["tf"], st["entry"], zeros, st["exit"], st["covered"], None, 1)
    pool_glsl, hash_glsl = _generated(tmp_path, s, pool, p)
    exe = glsl_ref.build_iso(tmp_path, pool_glsl, hash_glsl, color=True)
    u = orc.uniforms(p)
    g, hash_g = glsl_ref.run_iso(exe, tmp_path, p, u, orc.ray_exit_eye(p), st["entry"], zeros, st["covered"], st["meta"],
                                 pool.meta_dim, st["atlas"])
    assert not hash_g.any()
    hit_o, nrm_o = outs[0].reshape(-1, 4), outs[1].reshape(-1, 4)
    assert np.array_equal(hit_o[:, 3] != 0, g[0][:, 3] != 0) and (hit_o[:, 3] != 0).sum() > 200
    assert float(np.abs(hit_o[:, :3] - g[0][:, :3]).max()) <= 2e-5
    assert float(np.abs(nrm_o[:, :3] - g[1][:, :3]).max()) <= 2e-4
    # the colour packed into the two alpha channels: r + 1, floor(g * 512) + b
    assert float(np.abs(hit_o[:, 3] - g[0][:, 3]).max()) <= 1e-5
    hit = hit_o[:, 3] != 0
    assert (np.floor(nrm_o[hit, 3]) == np.floor(g[1][hit, 3])).mean() > 0.995 and (hit_o[hit, 3] > 1.0).any()
    assert np.array_equal(outs[2].reshape(-1, 4)[:, 3] == 1000.0, g[2][:, 3] == 1000.0)
    # Compose-Color-FS on the SAME buffers (the oracle's), executed vs restated
    cexe = glsl_ref.build_compose(tmp_path, color=True)
    amb = [p.ambient[i] * p.ambient[3] for i in range(3)]
    dif = [p.diffuse[i] * p.diffuse[3] for i in range(3)]             # no isosurface colour for colour data
    spe = [p.specular[i] * p.specular[3] for i in range(3)]
    img_g = glsl_ref.run_compose(cexe, tmp_path, s.width, s.height, amb, dif, spe, list(p.light_dir), outs[0], outs[1])
    img_o = orc.iso_compose(p, outs[0], outs[1]).reshape(-1, 4)
    assert float(np.abs(img_o - img_g).max()) <= 2e-4
    assert img_o[:, :3].std(axis=0).min() > 0.01                      # the surface carries the volume's colours


# ---------------------------------------------------------------------------------------------- the CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("mode,lighting", MODES + [(orc.RM_ISOSURFACE, False)])
def test_cuda_colour_frames_are_bit_identical_to_the_oracle(mode, lighting):
    s = scene(mode, lighting)
    want = s.oracle_render()
    r = s.make_renderer()
    st = r.PaintUntilConverged()
    assert st.converged
    got = r.ReadRGBA32F()
    assert np.array_equal(got, want["image"])
    assert np.array_equal(r.ReadRGBA8(), want["rgba8"])
    assert np.array_equal(r.page_table(), want["meta"])
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_colour_miss_lists_per_subframe_equal_the_oracle():
    s = scene(orc.RM_2DTRANS, True, size=(64, 48, 40))
    want = s.oracle_render()
    r = s.make_renderer()
    reqs = []
    for _ in range(64):
        st = r.Paint()
        reqs.append(r.missing_list())
        if st.converged:
            break
    assert len(reqs) == want["subframes"]
    for a, b in zip(reqs, want["requests"]):
        assert np.array_equal(np.asarray(a).reshape(-1, 4), np.asarray(b).reshape(-1, 4))
    assert np.array_equal(r.ReadRGBA32F(), want["image"])
    r.Cleanup()


@pytest.mark.gpu
def test_colour_volumes_are_refused_where_they_are_not_built():
    s = scene(orc.RM_1DTRANS, False)
    r = s.make_renderer()
    with pytest.raises(tb.TvkError):
        r.PaintClassic()               # the classic GLRaycaster path (GLRaycaster-Color-FS.glsl) is not built
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_colour_file_source_renders_like_the_registered_dataset():
    """A colour ExtendedOctree file on the streaming path (tvk_open_octree_file with the alpha min / max table, what
    tvk_open_uvf hands over from the MaxMin block): frames and page table identical to the same bricks served by callback."""
    path, (kind, size, dt, _, brick, ov, _, _), vol = _golden_colour_file()
    s = scene(orc.RM_2DTRANS, True)
    assert tuple(s.size) == tuple(size) and s.brick[0] == brick and np.array_equal(s.volume, vol)
    want = s.oracle_render()
    r = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    info = r.OpenOctreeFile(path, range_max=s.range_max, max_gradient_magnitude=s.max_grad)   # alpha min / max computed on the device
    assert info.dtype == tb.RGBA8
    n = r.info().total_bricks
    assert np.array_equal(r.minmax(n)[:, :2], s.octree.minmax[:n, :2])
    info = r.OpenOctreeFile(path, minmax=s.octree.minmax, range_max=s.range_max, max_gradient_magnitude=s.max_grad)
    assert info.dtype == tb.RGBA8
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
    r.Resize(s.width, s.height); r.SetRotation(s.rotation)
    r.CreateVolumePool(s._pool_size)
    assert r.PaintUntilConverged().converged
    assert np.array_equal(r.ReadRGBA32F(), want["image"])
    assert np.array_equal(r.page_table(), want["meta"])
    r.Cleanup()


@pytest.mark.gpu
def test_cuda_colour_uvf_opens_and_renders():
    """tvk_open_uvf on the colour golden container: TOC block + the MaxMin block's alpha component, then the colour kernels."""
    import os
    path, (kind, size, dt, _, brick, ov, _, _), vol = _golden_colour_file()
    uvf = os.path.join(os.path.dirname(path), "volume_rgba8_zlib.uvf")
    s = scene(orc.RM_1DTRANS, True)
    want = s.oracle_render()
    r = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    info = r.OpenUVF(uvf, range_max=s.range_max, max_gradient_magnitude=s.max_grad)
    assert info.dtype == tb.RGBA8
    n = r.info().total_bricks
    assert np.array_equal(r.minmax(n)[:, :2], s.octree.minmax[:n, :2])
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
    r.Resize(s.width, s.height); r.SetRotation(s.rotation)
    r.CreateVolumePool(s._pool_size)
    assert r.PaintUntilConverged().converged
    assert np.array_equal(r.ReadRGBA32F(), want["image"])
    r.Cleanup()


@pytest.mark.gpu
@pytest.mark.parametrize("size,brick", [((48, 40, 36), 16), ((33, 17, 41), 12), ((80, 72, 76), 36)])
def test_cuda_colour_bricker_matches_the_oracle(size, brick):
    """tvk_build_volume on a colour volume: per-component mean pyramid incl. the converter's odd-corner rule, 4-byte voxels
    cut by the float mover (36^3 bricks of an 80-voxel row: through the TMA box loads), alpha statistics from the store --
    every brick and every min / max equal to the oracle's ColorOctree (which is pinned to the unmodified converter), and
    the frame equal to the callback source's."""
    s = scene(orc.RM_2DTRANS, True, size=size, brick=brick)
    o = s.octree
    r = s.make_renderer("device")
    n_all = o.total_bricks
    mm = r.minmax(n_all)
    inner = brick - 4
    checked = 0
    for (x, y, z, lod) in o.iter_bricks():
        ls = o.lod_size(lod)
        if any(0 < (ls[a] % inner) < 2 and o.brick_count(lod)[a] > 1 for a in range(3)):
            continue                                   # "Q2": outside the oracle's contract, as for scalar volumes
        i = o.brick_index(x, y, z, lod)
        got = r.brick(x, y, z, lod, tb.RGBA8)
        assert np.array_equal(got.reshape(o.brick(x, y, z, lod).shape), o.brick(x, y, z, lod)), (x, y, z, lod)
        assert (mm[i, 0], mm[i, 1]) == (o.minmax[i, 0], o.minmax[i, 1]), (x, y, z, lod)
        checked += 1
    assert checked >= 3
    if size == (48, 40, 36):
        assert r.PaintUntilConverged().converged
        assert np.array_equal(r.ReadRGBA32F(), s.oracle_render()["image"])
    r.Cleanup()
