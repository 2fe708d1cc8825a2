"""Writes tests/golden/octree_*.bin: ExtendedOctree files produced by the UNMODIFIED reference converter
(oracle/_ref/ref_octree = ExtendedOctreeConverter compiled in place from /root/reference, with the reference's own
zlib / LZ4 / LZMA / bzip2 wrappers and vendored codecs) from the seeded synthetic volume below.  The reader tests (tests/test_octree_file.py) open these
on machines without the reference tree.  Run from the repo root:  python tests/golden/make_octree_golden.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tuvok_b200 import synth  # noqa: E402

CASES = {   # name: (kind, (x, y, z), dtype code, dtype name, brick, overlap, compression, layout)
    "octree_u16_none": (synth.V_NOISE, (26, 22, 18), 1, "u16", 12, 2, 0, 0),
    "octree_u16_lz4_morton": (synth.V_NOISE, (44, 36, 28), 1, "u16", 16, 2, 3, 1),
    "octree_u8_zlib_hilbert": (synth.V_SPH, (40, 40, 24), 0, "u8", 12, 2, 1, 2),
    "octree_f32_none": (synth.V_SPH, (18, 14, 12), 2, "f32", 12, 2, 0, 0),
    "octree_u16_lzma": (synth.V_NOISE, (40, 30, 26), 1, "u16", 16, 2, 2, 0),          # LZMA SDK (LzmaCompression.cpp)
    "octree_u8_bzip2_morton": (synth.V_SPH, (36, 32, 28), 0, "u8", 12, 2, 4, 1),      # bzip2 (BzlibCompression.cpp)
    # large bricks, to be re-cut on load (DynamicBrickingDS; tests/test_rebrick.py): ragged last source bricks on every axis
    "octree_u8_b36_zlib": (synth.V_SPH, (80, 70, 50), 0, "u8", 36, 2, 1, 0),
    "octree_u16_b28_lz4": (synth.V_NOISE, (60, 52, 30), 1, "u16", 28, 2, 3, 0),
    # a colour (RGBA8, four interleaved components) file: the multi-component path of the converter (tests/test_color.py)
    "octree_rgba8_zlib": (synth.V_SPH, (48, 40, 36), 3, "rgba8", 16, 2, 1, 0),
}


def volume(name):
    kind, size, dt, _, _, _, _, _ = CASES[name]
    if dt == 3:      # RGBA8: the alpha channel is the case's field, the colour channels three seeded noise fields
        import numpy as np
        chans = [synth.synth_volume(synth.V_NOISE, size, 0, 0x5EED + 11 * (k + 1)) for k in range(3)]
        return np.ascontiguousarray(np.stack(chans + [synth.synth_volume(kind, size, 0, 0x5EED)], axis=-1))
    return synth.synth_volume(kind, size, dt, 0x5EED)


# complete .uvf containers written by the reference's UVF / TOCBlock / Histogram1DDataBlock / MaxMinDataBlock /
# KeyValuePairDataBlock classes (oracle/_ref/ref_uvf): name: (volume case, compression, layout, timesteps)
UVF_CASES = {
    "volume_u8_zlib.uvf": ("octree_u8_zlib_hilbert", 1, 2, 1),
    "volume_rgba8_zlib.uvf": ("octree_rgba8_zlib", 1, 0, 1),        # colour: TOC block + four-component MaxMin block
}


if __name__ == "__main__":
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_octree")
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (kind, size, dt, dname, brick, ov, comp, layout) in CASES.items():
        raw = "/tmp/%s.raw" % name
        volume(name).tofile(raw)
        dst = os.path.join(out_dir, name + ".bin")
        subprocess.check_call([tool, raw, "/tmp/%s.dump" % name, dname, str(size[0]), str(size[1]), str(size[2]), str(brick),
                               str(ov), "0", "0", dst, str(comp), str(layout)], stdout=subprocess.DEVNULL)
        print(name, os.path.getsize(dst), "bytes")
    tool = os.path.join(ROOT, "oracle", "_ref", "ref_uvf")
    for name, (case, comp, layout, ts) in UVF_CASES.items():
        kind, size, dt, dname, brick, ov, _, _ = CASES[case]
        raw = "/tmp/%s.raw" % case
        volume(case).tofile(raw)
        dst = os.path.join(out_dir, name)
        subprocess.check_call([tool, raw, dst, dname, str(size[0]), str(size[1]), str(size[2]), str(brick), str(ov), str(comp),
                               str(layout), str(ts)], stdout=subprocess.DEVNULL)
        print(name, os.path.getsize(dst), "bytes")
