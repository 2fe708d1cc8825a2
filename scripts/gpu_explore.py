"""Ad-hoc GPU bring-up script (not a test): compares product and oracle over a grid of scenes and
prints the differences.  Run with `gpurun -- python tests/gpu_explore.py`."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scene import Scene, image_diff  # noqa: E402
from oracle import orc  # noqa: E402
import tuvok_b200 as tb  # noqa: E402
from tuvok_b200 import synth  # noqa: E402


def compare(name, s, source="device"):
    t0 = time.time()
    ref = s.oracle_render()
    t1 = time.time()
    r = s.make_renderer(source)
    r.enable_counters(True)
    st = r.PaintUntilConverged()
    img = r.ReadRGBA8()
    f32 = r.ReadRGBA32F()
    t2 = time.time()
    mx, psnr = image_diff(img, ref["rgba8"])
    fd = float(np.abs(f32 - ref["image"]).max())
    meta = r.page_table()
    table_ok = np.array_equal(meta, ref["meta"])
    mm_ok = True
    if source == "device":
        mm = r.minmax(len(s.octree.minmax))
        mm_ok = np.array_equal(mm, s.octree.minmax)
    print("%-28s max|d|=%d psnr=%6.1f fdiff=%.2e table=%s minmax=%s conv=%d paged=%d(ref %d) samples=%d(ref %d) "
          "bricks=%d(ref %d) oracle %.2fs gpu %.2fs" %
          (name, mx, psnr, fd, table_ok, mm_ok, st.converged, st.bricks_paged, ref["paged"], st.samples,
           ref["stats"].samples, st.brick_visits, ref["stats"].brick_visits, t1 - t0, t2 - t1), flush=True)
    if not table_ok:
        d = np.nonzero(meta != ref["meta"])[0]
        print("   table diff at", d[:10], meta[d[:10]], ref["meta"][d[:10]])
    if not mm_ok:
        d = np.nonzero((mm != s.octree.minmax).any(axis=1))[0]
        print("   minmax diff at", d[:10], mm[d[:3]], s.octree.minmax[d[:3]])
    r.Cleanup()
    return mx, psnr


def main():
    rot = (tb.rotation_y(30.0) @ tb.rotation_x(20.0)).astype(np.float32)
    compare("sph u8 1D single-brick", Scene(kind=synth.V_SPH, size=(32, 32, 32), dtype=orc.U8, brick=36, overlap=2,
                                           width=64, height=64, tf_center=0.3, tf_inv_gradient=0.3))
    compare("sph u8 1D bricked", Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2,
                                      width=96, height=96, tf_center=0.3, tf_inv_gradient=0.3))
    compare("sph u8 1D bricked cb", Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2,
                                         width=96, height=96, tf_center=0.3, tf_inv_gradient=0.3), "callback")
    for mode, lit, name in [(orc.RM_1DTRANS, False, "1D"), (orc.RM_1DTRANS, True, "1D-L"), (orc.RM_2DTRANS, False, "2D"),
                            (orc.RM_2DTRANS, True, "2D-L"), (orc.RM_ISOSURFACE, True, "iso")]:
        for dt, dn in [(orc.U8, "u8"), (orc.U16, "u16"), (orc.F32, "f32")]:
            iso = {orc.U8: 80, orc.U16: 20000, orc.F32: 0.3}[dt]
            s = Scene(kind=synth.V_NOISE, size=(72, 64, 56), dtype=dt, brick=20, overlap=2, width=128, height=96,
                      mode=mode, lighting=lit, rotation=rot, tf_center=0.3, tf_inv_gradient=0.3, isovalue=iso)
            compare("noise %s %s rot" % (dn, name), s)
    s = Scene(kind=synth.V_NOISE, size=(96, 96, 96), dtype=orc.U16, brick=20, overlap=2, width=160, height=120,
              mode=orc.RM_2DTRANS, lighting=True, rotation=rot, translation=tb.translation(0.1, 0, 1.2),
              tf_center=0.3, tf_inv_gradient=0.3)
    compare("noise u16 2D-L inside", s)
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U16, brick=20, overlap=2, width=96, height=96,
              mode=orc.RM_1DTRANS, lighting=True, rotation=rot, sample_rate=2.0, tf_center=0.3, tf_inv_gradient=0.3)
    compare("sph u16 1D-L rate2", s)
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, width=96, height=96,
              nearest=True, lighting=True, rotation=rot, tf_center=0.3, tf_inv_gradient=0.3)
    compare("sph u8 1D-L nearest", s)
    # small pool: eviction / LRU
    s = Scene(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, width=96, height=96,
              rotation=rot, pool_size=(60, 60, 40), tf_center=0.3, tf_inv_gradient=0.3)
    compare("sph u8 1D small pool", s)


if __name__ == "__main__":
    main()
