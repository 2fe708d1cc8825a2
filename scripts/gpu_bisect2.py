import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_scenes
from oracle import orc
from scene import Scene, image_diff
from tuvok_b200 import synth
for kw in (dict(nearest=True, lighting=True), dict(nearest=True), dict(overlap=1, brick=18, lighting=True), dict(overlap=1, brick=18),
           dict(overlap=1, brick=18, mode=orc.RM_2DTRANS)):
    base = dict(kind=synth.V_SPH, size=(64, 64, 64), dtype=orc.U8, brick=20, overlap=2, width=96, height=96,
                rotation=golden_scenes.ROT, tf_center=0.3, tf_inv_gradient=0.3)
    base.update(kw)
    s = Scene(**base)
    ref = s.oracle_render()
    r = s.make_renderer("device")
    r.enable_counters(True)
    st = r.PaintUntilConverged()
    img = r.ReadRGBA8()
    mx, psnr = image_diff(img, ref["rgba8"])
    d = np.abs(img.astype(int) - ref["rgba8"].astype(int)).max(axis=2)
    print(kw, "max", mx, "psnr %.1f" % psnr, "bad px", int((d > 2).sum()), "samples", st.samples, ref["stats"].samples, "table", np.array_equal(r.page_table(), ref["meta"]), flush=True)
    ys, xs = np.nonzero(d > 2)
    for y, x in list(zip(ys, xs))[:4]:
        print("   ", x, y, img[y, x], ref["rgba8"][y, x])
    r.Cleanup()
