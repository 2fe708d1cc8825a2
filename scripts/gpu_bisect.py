"""Developer tool (under gpurun): localise float-level differences between the CUDA renderer and the live oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_scenes
from oracle import orc

def run(tag, **over):
    s = golden_scenes.make("ragged_1d_lit", **over)
    ref = s.oracle_render()
    r = s.make_renderer("device")
    st = r.PaintUntilConverged()
    f = r.ReadRGBA32F()
    d = np.abs(f - ref["image"])
    print("%-30s subframes %d max|df| %.3e bad px %d  (alpha diff %.3e)" % (tag, ref["subframes"], d.max(), int((d.max(axis=2) > 0).sum()), d[..., 3].max()), flush=True)
    r.Cleanup()

run("as is")
run("lighting off", lighting=False)
run("big pool, same", max_gpu_mem=4 << 30)
run("cubic 64", size=(64, 64, 64))
run("size 70,45,58 w 96x96", width=96, height=96)
run("brick 36", brick=36)
run("u16", dtype=orc.U16)
run("no rotation", rotation=np.eye(4, dtype=np.float32))
run("tf .3/.3", tf_center=0.3, tf_inv_gradient=0.3)
