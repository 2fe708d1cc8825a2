"""Developer tool (run under gpurun): float-level diff of golden scenes between the CUDA renderer and the golden
vectors.  python scripts/gpu_diff.py [scene ...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_scenes
names = sys.argv[1:] or sorted(golden_scenes.SCENES)
for name in names:
    for source in ("device", "callback"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        s = golden_scenes.make(name)
        r = s.make_renderer(source)
        r.enable_counters(True)
        st = r.PaintUntilConverged()
        f = r.ReadRGBA32F(); ref = g["image"]
        d = np.abs(f - ref)
        bad = np.argwhere(d.max(axis=2) > 0)
        print("%-24s %-8s conv=%d samples=%d (gold %d) max|df|=%.3e n_bad_px=%d" %
              (name, source, st.converged, st.samples, int(g["samples"]), d.max(), len(bad)), flush=True)
        for y, x in bad[:6]:
            print("    px (%d,%d) gpu %s gold %s" % (x, y, f[y, x], ref[y, x]))
        r.Cleanup()
