"""ExtendedOctree file reader (tuvok_b200/csrc/octree_file.cpp behind tvk_open_octree_file).

CPU (host-only C ABI helpers, no device):
  * golden files written by the UNMODIFIED reference converter (tests/golden/octree_*.bin, made by
    tests/golden/make_octree_golden.py through oracle/_ref/ref_octree; uncompressed, zlib, LZ4; scanline / Morton /
    Hilbert layouts): header, TOC and every decoded brick == the oracle's bricker on the same seeded volume;
  * fresh files from the reference converter when oracle/_ref/ref_octree is present (random volumes);
  * malformed files are refused with an error, never read out of bounds.
GPU: the file as brick source of the streaming path -- device-computed min/max table == oracle, bricks through the
pool bit-exact, rendered image bit-identical to the same scene from the GPU bricker."""
import os
import subprocess

import numpy as np
import pytest

import tuvok_b200 as tb
from oracle import orc
from tuvok_b200 import _lib as L
from tuvok_b200 import octree_file, synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("make_octree_golden", os.path.join(GOLDEN, "make_octree_golden.py"))
golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(golden)


def check_file(path, vol, brick, overlap, dtype, codec_slot=None):
    # (a colour file: four interleaved components, the oracle's ColorOctree)
    o = orc.ColorOctree(vol, brick, overlap) if dtype == orc.RGBA8 else orc.Octree(vol, brick, overlap)
    info = octree_file.probe(path)
    assert tuple(info.domain_size) == (vol.shape[2], vol.shape[1], vol.shape[0])
    assert tuple(info.max_brick_size) == (brick,) * 3 and info.overlap == overlap and info.dtype == dtype
    assert info.lod_count == o.lod_count and info.brick_count == o.total_bricks and info.version >= 1
    if codec_slot is not None:
        assert info.bricks_by_codec[codec_slot] > 0
    assert sum(info.bricks_by_codec) == o.total_bricks
    inner = brick - 2 * overlap
    for (x, y, z, lod) in o.iter_bricks():
        got = octree_file.read_brick(path, x, y, z, lod, info=info)
        want = o.brick(x, y, z, lod)
        assert got.shape == want.shape, (x, y, z, lod)
        # outside the oracle's contract (orc_octree.c header, "Q2"; same exclusion as tests/test_octree_ref.py): a last
        # brick whose remainder is smaller than the overlap makes the reference CONVERTER read stale memory
        ls = o.lod_size(lod)
        if any(0 < (ls[a] % inner) < overlap and o.brick_count(lod)[a] > 1 for a in range(3)):
            continue
        assert np.array_equal(got, want), (x, y, z, lod)
    return o, info


@pytest.mark.parametrize("name", sorted(golden.CASES))
def test_golden_reference_written_files(name):
    kind, size, dt, dname, brick, ov, comp, layout = golden.CASES[name]
    slot = {0: 0, 1: 1, 2: 2, 3: 3, 4: 4}[comp]
    _, info = check_file(os.path.join(GOLDEN, name + ".bin"), golden.volume(name), brick, ov, dt, codec_slot=slot)
    if comp:
        # a compressor may store single bricks raw when they do not shrink; most must be compressed
        assert info.bricks_by_codec[slot] >= info.brick_count // 2
        assert info.payload_bytes < sum(b.nbytes for b in [golden.volume(name)]) * 3


@pytest.mark.parametrize("shape,dtype,brick,overlap,comp,layout", [
    ((33, 17, 40), orc.U16, 12, 2, 3, 0),
    ((30, 26, 22), orc.U8, 14, 3, 1, 1),
    ((36, 36, 36), orc.U16, 10, 1, 0, 3),      # random brick order on disk
    ((32, 32, 32), orc.F32, 12, 2, 3, 2),
    ((37, 29, 41), orc.U16, 14, 2, 2, 0),      # LZMA
    ((32, 32, 32), orc.F32, 12, 2, 2, 1),      # LZMA, floats, Morton order
    ((30, 26, 22), orc.U8, 14, 3, 4, 2),       # bzip2
    ((64, 64, 64), orc.U16, 36, 2, 2, 0),      # LZMA, 36^3 bricks (93 KB each: long matches, many rep codes)
])
def test_fresh_reference_files(ref_octree_bin, tmp_path, shape, dtype, brick, overlap, comp, layout):
    rng = np.random.default_rng(sum(shape) + comp)
    if dtype == orc.F32:
        vol = np.round(rng.random(shape, dtype=np.float32) * 8) / 8       # compressible floats
    else:
        vol = (rng.integers(0, 6, size=shape) * (9 if dtype == orc.U8 else 2000)).astype(orc.NP_DTYPE[dtype])
    raw = tmp_path / "in.raw"
    vol.tofile(raw)
    dst = tmp_path / "octree.bin"
    name = {orc.U8: "u8", orc.U16: "u16", orc.F32: "f32"}[dtype]
    subprocess.check_call([ref_octree_bin, str(raw), str(tmp_path / "dump.bin"), name, str(shape[2]), str(shape[1]),
                           str(shape[0]), str(brick), str(overlap), "0", "0", str(dst), str(comp), str(layout)],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    check_file(str(dst), vol, brick, overlap, dtype)


def test_embedded_at_an_offset(tmp_path):
    """Inside a UVF the octree is the payload of a TOC block: header at a byte offset, TOC offsets relative to it."""
    src = open(os.path.join(GOLDEN, "octree_u16_lz4_morton.bin"), "rb").read()
    p = tmp_path / "embedded.bin"
    p.write_bytes(b"UVF-DATA" + bytes(1234) + src + bytes(77))
    info = octree_file.probe(str(p), offset=1242)
    kind, size, dt, _, brick, ov, _, _ = golden.CASES["octree_u16_lz4_morton"]
    assert tuple(info.domain_size) == size
    o = orc.Octree(golden.volume("octree_u16_lz4_morton"), brick, ov)
    for key in [(0, 0, 0, 0), (2, 1, 1, 0), (0, 0, 0, o.lod_count - 1)]:
        got = octree_file.read_brick(str(p), *key, info=info, offset=1242)
        assert np.array_equal(got, o.brick(*key))


def test_malformed_files_are_refused(tmp_path):
    src = bytearray(open(os.path.join(GOLDEN, "octree_u8_zlib_hilbert.bin"), "rb").read())
    with pytest.raises(L.TvkError):
        octree_file.probe(str(tmp_path / "missing.bin"))
    t = tmp_path / "trunc.bin"
    t.write_bytes(bytes(src[:60]))
    with pytest.raises(L.TvkError):
        octree_file.probe(str(t))
    t = tmp_path / "short.bin"
    t.write_bytes(bytes(src[:len(src) - 500]))          # last bricks beyond the end of the file
    with pytest.raises(L.TvkError):
        octree_file.probe(str(t))
    z = bytearray(src)
    z[12 + 1:12 + 1 + 8] = bytes(8)                       # volume size x = 0
    t = tmp_path / "zero.bin"
    t.write_bytes(bytes(z))
    with pytest.raises(L.TvkError):
        octree_file.probe(str(t))
    info = octree_file.probe(os.path.join(GOLDEN, "octree_u8_zlib_hilbert.bin"))
    c = bytearray(src)
    c[-400:-300] = bytes(range(100))                      # corrupt a compressed stream
    t = tmp_path / "corrupt.bin"
    t.write_bytes(bytes(c))
    bad = 0
    kind, size, dt, _, brick, ov, _, _ = golden.CASES["octree_u8_zlib_hilbert"]
    o = orc.Octree(golden.volume("octree_u8_zlib_hilbert"), brick, ov)
    for key in o.iter_bricks():
        try:
            if not np.array_equal(octree_file.read_brick(str(t), *key, info=info), o.brick(*key)):
                bad += 1            # (a stored-raw brick: the damage shows up as wrong voxels instead)
        except L.TvkError:
            bad += 1
    assert bad >= 1
    with pytest.raises(L.TvkError):
        octree_file.read_brick(os.path.join(GOLDEN, "octree_u8_zlib_hilbert.bin"), 99, 0, 0, 0, info=info)


@pytest.mark.parametrize("name", ["octree_u16_lzma", "octree_u8_bzip2_morton"])
def test_corrupt_lzma_and_bzip2_streams_never_pass_silently(tmp_path, name):
    """Damage inside a compressed brick: the decoder reports it or the voxels differ -- and nothing is written past the brick."""
    src = bytearray(open(os.path.join(GOLDEN, name + ".bin"), "rb").read())
    kind, size, dt, _, brick, ov, comp, _ = golden.CASES[name]
    info = octree_file.probe(os.path.join(GOLDEN, name + ".bin"))
    o = orc.Octree(golden.volume(name), brick, ov)
    rng = np.random.default_rng(11)
    for trial in range(6):
        c = bytearray(src)
        at = int(rng.integers(len(c) // 2, len(c) - 64))
        c[at:at + 24] = bytes(int(v) for v in rng.integers(0, 256, 24))
        t = tmp_path / ("corrupt%d.bin" % trial)
        t.write_bytes(bytes(c))
        bad = 0
        for key in o.iter_bricks():
            try:
                if not np.array_equal(octree_file.read_brick(str(t), *key, info=info), o.brick(*key)):
                    bad += 1
            except L.TvkError:
                bad += 1
        assert bad >= 1, (trial, at)


def test_lz4_decoder_against_python_roundtrip(tmp_path):
    """Literal-only and long-match blocks: the decoder on hand-built LZ4 blocks (format: token, literals, offset,
    extended lengths) -- exercised through a synthetic single-brick octree file written here."""
    import struct
    n = 8
    vox = np.zeros((n, n, n), np.uint8)
    vox[2:6, 2:6, 2:6] = 200
    raw = vox.tobytes()

    def lz4_block(data):          # greedy RLE-ish encoder producing valid LZ4 sequences (tests the decoder only)
        out, i, lit = bytearray(), 0, bytearray()
        while i < len(data):
            j = i
            while j < len(data) and data[j] == data[i]:
                j += 1
            run = j - i
            if run >= 8 and j < len(data):     # emit literals + first byte, then a match (offset 1) for the rest
                lit.append(data[i])
                ml = run - 1
                tok_l, tok_m = min(len(lit), 15), min(ml - 4, 15)
                out.append(tok_l << 4 | tok_m)
                if tok_l == 15:
                    r = len(lit) - 15
                    while r >= 255:
                        out.append(255); r -= 255
                    out.append(r)
                out += lit
                out += struct.pack("<H", 1)
                if tok_m == 15:
                    r = ml - 4 - 15
                    while r >= 255:
                        out.append(255); r -= 255
                    out.append(r)
                lit = bytearray()
                i = j
            else:
                lit += data[i:j]
                i = j
        tok_l = min(len(lit), 15)
        out.append(tok_l << 4)
        if tok_l == 15:
            r = len(lit) - 15
            while r >= 255:
                out.append(255); r -= 255
            out.append(r)
        out += lit
        return bytes(out)

    blk = lz4_block(raw)
    assert len(blk) < len(raw)
    # octree header (version 2): one 8^3 brick (brick 8 = volume 4 + 2*2 ghost), LoDs 4 -> 2 -> 1: three bricks
    hdr = struct.pack("<IQ?QQQdddQQQIIQI", 0, 1, False, 4, 4, 4, 1.0, 1.0, 1.0, 8, 8, 8, 2, 2, 0, 1)
    sizes = [8, 6, 5]
    toc_len = 3 * 36
    body, toc, off = b"", b"", len(hdr) + toc_len
    bricks = []
    for s in sizes:
        b = np.zeros((s, s, s), np.uint8)
        b[1:-1, 1:-1, 1:-1] = 7
        if s == 8:
            b, payload, codec = vox, blk, 3
        else:
            payload, codec = b.tobytes(), 0
        bricks.append(b)
        toc += struct.pack("<QQIQII", off, len(payload), codec, len(payload), 0, 0)
        body += payload
        off += len(payload)
    p = tmp_path / "hand.bin"
    p.write_bytes(hdr + toc + body)
    info = octree_file.probe(str(p))
    assert info.lod_count == 3 and info.brick_count == 3 and info.bricks_by_codec[3] == 1
    for lod, b in enumerate(bricks):
        assert np.array_equal(octree_file.read_brick(str(p), 0, 0, 0, lod, info=info), b)


# ------------------------------------------------------------------------------------------------ UVF container
def check_uvf(path, vol, brick, overlap, dtype, timesteps=1):
    o = orc.Octree(vol, brick, overlap)
    for ts in range(timesteps):
        u = octree_file.uvf_probe(path, ts)
        assert u["file_version"] == 5 and u["n_timesteps"] == timesteps
        assert u["n_blocks"] >= 2 * timesteps + 1                 # TOC + MaxMin per timestep (+ histogram), key/value block
        # the MaxMin block is the oracle's (= the reference converter's BrickStatVec) min/max table, TOC order
        assert u["maxmin"].shape == (o.total_bricks, 4)
        assert np.array_equal(u["maxmin"], o.minmax)
        info = octree_file.probe(path, offset=u["toc_payload_offset"], uvf_file_version=u["file_version"])
        assert info.brick_count == o.total_bricks and info.dtype == dtype
        for key in list(o.iter_bricks())[::3]:
            got = octree_file.read_brick(path, *key, info=info, offset=u["toc_payload_offset"])
            assert np.array_equal(got, o.brick(*key)), key
    with pytest.raises(L.TvkError):
        octree_file.uvf_probe(path, timesteps)                    # timestep out of range


def test_golden_uvf_container():
    case, comp, layout, ts = golden.UVF_CASES["volume_u8_zlib.uvf"]
    kind, size, dt, _, brick, ov, _, _ = golden.CASES[case]
    check_uvf(os.path.join(GOLDEN, "volume_u8_zlib.uvf"), golden.volume(case), brick, ov, dt, ts)


def test_golden_uvf_stats():
    """The container's other blocks (machines without the reference tree): range from the MaxMin block, 1D histogram."""
    st = octree_file.uvf_stats(os.path.join(GOLDEN, "volume_u8_zlib.uvf"))
    vol = golden.volume("octree_u8_zlib_hilbert")
    assert st["range"] == (float(vol.min()), float(vol.max()))
    assert st["hist1d_filled"] == int(vol.max()) + 1
    with pytest.raises(L.TvkError):
        octree_file.uvf_stats(os.path.join(GOLDEN, "octree_u16_none.bin"))


def test_fresh_uvf_with_two_timesteps(tmp_path):
    tool = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_uvf")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/ref_uvf not built (reference tree absent)")
    vol = synth.synth_volume(synth.V_NOISE, (44, 36, 28), orc.U16, 0x5EED)
    raw = tmp_path / "in.raw"
    vol.tofile(raw)
    dst = tmp_path / "two.uvf"
    subprocess.check_call([tool, str(raw), str(dst), "u16", "44", "36", "28", "16", "2", "3", "1", "2"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    check_uvf(str(dst), vol, 16, 2, orc.U16, timesteps=2)


def test_uvf_walk_refuses_garbage(tmp_path):
    src = open(os.path.join(GOLDEN, "volume_u8_zlib.uvf"), "rb").read()
    for name, data in [("magic", b"UVF-DATB" + src[8:]), ("trunc", src[:70]), ("octree_only", open(os.path.join(
            GOLDEN, "octree_u16_none.bin"), "rb").read())]:
        p = tmp_path / (name + ".uvf")
        p.write_bytes(data)
        with pytest.raises(L.TvkError):
            octree_file.uvf_probe(str(p))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["octree_u16_lz4_morton", "octree_u8_zlib_hilbert", "octree_f32_none"])
def test_file_source_streams_into_the_pool(name):
    from scene import Scene, image_diff
    kind, size, dt, _, brick, ov, _, _ = golden.CASES[name]
    path = os.path.join(GOLDEN, name + ".bin")
    s = Scene(kind=kind, size=size, dtype=dt, brick=brick, overlap=ov, width=72, height=56,
              rotation=(tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32), lighting=dt != orc.F32,
              tf_center=0.25, tf_inv_gradient=0.3, seed=0x5EED)
    o = s.octree
    dev = s.make_renderer("device")              # the same scene bricked on the GPU
    dev.PaintUntilConverged()
    want = dev.ReadRGBA32F()

    r = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    info = r.OpenOctreeFile(path, max_gradient_magnitude=s.max_grad)    # min/max computed on the device
    assert info.brick_count == o.total_bricks
    mm = r.minmax(r.info().total_bricks)
    assert np.array_equal(mm, o.minmax[:len(mm)])
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
    r.SetIsoValue(s.isovalue); r.Resize(s.width, s.height); r.SetRotation(s.rotation)
    r.CreateVolumePool(s._pool_size)
    st = r.PaintUntilConverged()
    assert st.converged and st.bricks_paged > 0
    got = r.ReadRGBA32F()
    assert np.array_equal(got, want)             # bit-identical to the GPU-bricked scene
    assert np.array_equal(r.page_table(), dev.page_table())
    # and with the table handed in (MaxMinDataBlock), identical again
    r2 = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    r2.OpenOctreeFile(path, minmax=o.minmax, max_gradient_magnitude=s.max_grad)
    assert np.array_equal(r2.minmax(len(mm)), mm)
    for rr in (r, r2, dev):
        rr.Cleanup()


@pytest.mark.gpu
def test_uvf_file_renders_like_the_gpu_bricked_volume():
    from scene import Scene
    case, comp, layout, ts = golden.UVF_CASES["volume_u8_zlib.uvf"]
    kind, size, dt, _, brick, ov, _, _ = golden.CASES[case]
    s = Scene(kind=kind, size=size, dtype=dt, brick=brick, overlap=ov, width=72, height=56, lighting=True,
              rotation=(tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32), tf_center=0.25,
              tf_inv_gradient=0.3, seed=0x5EED)
    dev = s.make_renderer("device")
    dev.PaintUntilConverged()
    r = tb.CudaGridLeaper(max_gpu_mem=s.max_gpu_mem, hash_table_size=s.hash_size(), brick_strategy=s.strategy)
    # (range_max given: left at 0, tvk_open_uvf would take the file's own value range, 0..249, as UVFDataset does)
    info = r.OpenUVF(os.path.join(GOLDEN, "volume_u8_zlib.uvf"), range_max=s.range_max, max_gradient_magnitude=s.max_grad)
    assert info.brick_count == s.octree.total_bricks and info.bricks_by_codec[1] > 0
    n = r.info().total_bricks
    assert np.array_equal(r.minmax(n), s.octree.minmax[:n])          # from the file's MaxMin block
    r.Set1DTrans(s.tf1d); r.Set2DTrans(s.tf2d); r.SetRendermode(s.mode); r.SetUseLighting(s.lighting)
    r.Resize(s.width, s.height); r.SetRotation(s.rotation)
    r.CreateVolumePool(s._pool_size)
    assert r.PaintUntilConverged().converged
    assert np.array_equal(r.ReadRGBA32F(), dev.ReadRGBA32F())
    assert np.array_equal(r.page_table(), dev.page_table())
    r.Cleanup(); dev.Cleanup()
