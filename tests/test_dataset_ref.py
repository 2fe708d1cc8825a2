"""The data interface the renderers consume (SURVEY 8b "Data interfaces consumed") against the UNMODIFIED reference
UVFDataset: a .uvf written by the reference's own UVF classes (oracle/_ref/ref_uvf) is opened by the reference's own
UVFDataset (oracle/_ref/ref_dataset: IO/uvfDataset.cpp + BrickedDataset / LinearIndexDataset / TOCBlock / ...
compiled in place) and everything a renderer reads from it is dumped:
  * LoD count, GetLargestSingleBrickLOD, domain size and brick layout per LoD, overlap, max used brick size,
  * per brick: voxel counts, MaxMinForKey, the voxels GetBrick returns,
  * per brick: centre, extents and texture coordinates (BrickMD, GetTextCoords) -- the geometry of the classic
    per-brick path (a13/a14).
Compared bit for bit with the oracle (orc_octree.c geometry / bricks / min-max, orc_classic.cpp brick metadata) and
with the product's file source (tvk_uvf_probe + tvk_octree_file_probe)."""
import os
import subprocess

import numpy as np
import pytest

import tuvok_b200 as tb
from oracle import orc
from scene import Scene
from tuvok_b200 import octree_file, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_UVF = os.path.join(ROOT, "oracle", "_ref", "ref_uvf")
REF_DS = os.path.join(ROOT, "oracle", "_ref", "ref_dataset")


def _have():
    if not (os.path.exists(REF_UVF) and os.path.exists(REF_DS)) and os.path.isdir("/root/reference/IO"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(REF_UVF) and os.path.exists(REF_DS)


pytestmark = pytest.mark.skipif(not _have(), reason="oracle/_ref/ref_uvf / ref_dataset not built (reference tree absent)")


def fnv1a(data):
    h = 1469598103934665603
    for b in data:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def hx(tokens):
    return np.array([float.fromhex(t) for t in tokens], np.float64)


def dump(tmp_path, vol, dtype, brick, overlap, comp=0, layout=0):
    raw = tmp_path / "in.raw"
    vol.tofile(raw)
    uvf = tmp_path / "vol.uvf"
    name = {orc.U8: "u8", orc.U16: "u16"}[dtype]
    nz, ny, nx = vol.shape
    subprocess.check_call([REF_UVF, str(raw), str(uvf), name, str(nx), str(ny), str(nz), str(brick), str(overlap), str(comp),
                           str(layout), "1"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = tmp_path / "ds.txt"
    subprocess.check_call([REF_DS, str(uvf), str(out), "256"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    head, lods, bricks = {}, {}, {}
    for line in open(out):
        t = line.split()
        if t[0] == "lods":
            head.update(lods=int(t[1]), largest_single=int(t[3]), bits=int(t[5]), overlap=tuple(int(v) for v in t[13:16]),
                        maxused=tuple(int(v) for v in t[17:20]))
        elif t[0] == "scale":
            head.update(scale=hx(t[1:4]), range=hx(t[5:7]), maxgrad=float.fromhex(t[8]), total=int(t[10]))
        elif t[0] == "lod":
            lods[int(t[1])] = dict(domain=tuple(int(v) for v in t[3:6]), layout=tuple(int(v) for v in t[7:10]))
        elif t[0] == "brick":
            bricks[(int(t[1]), int(t[2]))] = dict(center=hx(t[4:7]), ext=hx(t[8:11]), vox=tuple(int(v) for v in t[12:15]),
                                                  tmin=hx(t[16:19]), tmax=hx(t[20:23]), mm=hx(t[24:26]),
                                                  first=tuple(int(v) for v in t[27:30]), last=tuple(int(v) for v in t[31:34]),
                                                  fnv=int(t[35], 16))
    return str(uvf), head, lods, bricks


CASES = [
    (synth.V_NOISE, (44, 36, 28), orc.U8, 16, 2, (1.0, 1.0, 1.0)),
    (synth.V_SPH, (40, 40, 24), orc.U8, 12, 2, (1.0, 1.0, 1.0)),
    (synth.V_NOISE, (70, 45, 58), orc.U8, 20, 2, (1.0, 1.0, 1.0)),      # ragged last bricks, odd LoD sizes (aspect != 1)
]


@pytest.mark.parametrize("kind,size,dtype,brick,overlap,scale", CASES)
def test_oracle_and_file_source_match_reference_uvfdataset(tmp_path, kind, size, dtype, brick, overlap, scale):
    vol = synth.synth_volume(kind, size, dtype, 0x5EED)
    uvf, head, lods, bricks = dump(tmp_path, vol, dtype, brick, overlap, comp=1)
    o = orc.Octree(vol, brick, overlap)
    # ---- geometry + data: the oracle's octree
    assert head["lods"] == o.lod_count and head["largest_single"] == o.largest_single_brick_lod
    assert head["overlap"] == (overlap,) * 3 and head["total"] == o.total_bricks
    inner = brick - 2 * overlap
    for lod in range(o.lod_count):
        assert lods[lod]["domain"] == tuple(o.lod_size(lod)) and lods[lod]["layout"] == tuple(o.brick_count(lod))
    q2 = {lod for lod in range(o.lod_count)
          if any(0 < (o.lod_size(lod)[a] % inner) < overlap and o.brick_count(lod)[a] > 1 for a in range(3))}
    for (x, y, z, lod) in o.iter_bricks():
        bc = o.brick_count(lod)
        b = bricks[(lod, x + y * bc[0] + z * bc[0] * bc[1])]
        assert b["vox"] == tuple(o.brick_size(x, y, z, lod))
        # (BrickIsFirst/LastInDimension compare centres across ALL LoDs, BrickedDataset.cpp:147-173, and are only used by
        #  the non-TOC Dataset::GetTextCoords; TOC datasets use UVFDataset::GetTextCoords, checked below)
        if lod in q2:
            continue          # outside the oracle's contract (the converter reads stale memory there, orc_octree.c "Q2")
        i = o.brick_index(x, y, z, lod)
        assert tuple(b["mm"]) == (o.minmax[i, 0], o.minmax[i, 1])
        assert b["fnv"] == fnv1a(np.ascontiguousarray(o.brick(x, y, z, lod)).tobytes())
    assert head["maxused"] == tuple(max(bricks[k]["vox"][a] for k in bricks) for a in range(3))
    # ---- the product's file source sees the same dataset
    u = octree_file.uvf_probe(uvf)
    info = octree_file.probe(uvf, offset=u["toc_payload_offset"], uvf_file_version=u["file_version"])
    assert info.lod_count == head["lods"] and info.brick_count == head["total"] and info.overlap == overlap
    assert tuple(info.domain_size) == lods[0]["domain"]
    assert np.array_equal(np.array(list(info.aspect)), head["scale"])      # UVFDataset::GetScale == the octree's aspect
    # ---- classic-path brick metadata (centre, extents, texture coordinates) of every pool LoD: orc_classic.cpp
    s = Scene(kind=kind, size=size, dtype=dtype, brick=brick, overlap=overlap, width=64, height=64, scale=scale,
              translation=tb.translation(0, 0, -40.0), seed=0x5EED)
    pool, _ = s.oracle_pool()
    p = s.oracle_params(pool)
    for lod in range(s.pool_lod_count()):
        if lod in q2:
            continue
        bc = o.brick_count(lod)
        n = bc[0] * bc[1] * bc[2]
        first = o.brick_index(0, 0, 0, lod)
        mine, cnt = orc.classic_brick_list(p, lod, overlap, o.minmax[first:first + n], (0.0, 1e30, 0.0, 1e30))
        assert cnt == n
        for j in range(cnt):
            m, r = mine[j], bricks[(lod, mine[j].index)]
            assert tuple(np.float32(v) for v in m.center) == tuple(np.float32(v) for v in r["center"]), (lod, m.index)
            assert tuple(np.float32(v) for v in m.ext) == tuple(np.float32(v) for v in r["ext"]), (lod, m.index)
            assert tuple(np.float32(v) for v in m.tex_min) == tuple(np.float32(v) for v in r["tmin"]), (lod, m.index)
            assert tuple(np.float32(v) for v in m.tex_max) == tuple(np.float32(v) for v in r["tmax"]), (lod, m.index)
            assert tuple(m.n_vox) == r["vox"]


@pytest.mark.parametrize("dtype,shape,brick,overlap", [(orc.U8, (40, 36, 28), 16, 2), (orc.U16, (33, 47, 52), 20, 2)])
def test_uvf_stats_match_what_uvfdataset_derives(tmp_path, dtype, shape, brick, overlap):
    """tvk_uvf_probe_stats / the defaults of tvk_open_uvf: value range == UVFDataset::GetRange (ComputeRange over the
    LoD-0 bricks of the MaxMin block), maximum gradient magnitude == UVFDataset::GetMaxGradMagnitude (2D histogram
    block), 1D histogram filled size == index of the last non-zero bin + 1."""
    from tuvok_b200 import octree_file
    rng = np.random.default_rng(3)
    top = 200 if dtype == orc.U8 else 3000
    vol = (rng.integers(0, top, size=shape)).astype(orc.NP_DTYPE[dtype])
    vol[5:20, 5:20, 5:20] //= 4                                   # some structure for the gradient histogram
    uvf, head, _, _ = dump(tmp_path, vol, dtype, brick, overlap)
    st = octree_file.uvf_stats(uvf)
    assert st["range"] == (float(head["range"][0]), float(head["range"][1]))
    assert st["range"] == (float(vol.min()), float(vol.max()))
    assert np.float32(st["max_gradient_magnitude"]) == np.float32(head["maxgrad"]) and st["max_gradient_magnitude"] > 0
    assert st["hist1d_filled"] == int(vol.max()) + 1 and st["hist1d_size"] >= st["hist1d_filled"]
    assert st["hist2d_size"][0] > 0 and st["hist2d_size"][1] > 0
