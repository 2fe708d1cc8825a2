"""Developer tool (one GPU): the CUDA depth-pipeline stages (tvk_render_stage, run one after the other on one renderer)
against the oracle's stages (Scene.oracle_pipeline) -- per stage the hand-over images and the accumulated colour."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import golden_scenes
from tuvok_b200 import sortlast
from test_sortlast import scene_layout

for name, n in [("c2_bricked36_1d_ert", 2), ("c3_bricked36_2d_lit", 2), ("ragged_1d_lit", 3), ("c2_bricked36_1d_ert", 3)]:
    s = golden_scenes.make(name)
    ref = s.oracle_pipeline(n, fresh=True)     # the renderer below only ever runs stages: same paging history
    finest, fl, ext = scene_layout(s)
    ren = s.make_renderer("device")
    n_pix = s.width * s.height
    pos = col = None
    rows = []
    for st_i, stage in enumerate(ref["stages"]):
        ren.SetShardBox(*sortlast.box_to_clip(stage["box"], finest, fl))
        for _ in range(32):
            st = ren.RenderStage(pos.data_ptr() if pos is not None else 0, col.data_ptr() if col is not None else 0)
            if st.converged:
                break
        ptrs = ren.stage_output_ptrs()

        def grab(ptr):
            class _Dev:
                __cuda_array_interface__ = {"shape": (n_pix, 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
            return torch.as_tensor(_Dev(), device="cuda").clone()
        img, col, pos = (grab(p) for p in ptrs)
        o = stage["outs"]
        rows.append("stage %d: image %s (max|d| %.3g) colour %s position %s" % (
            st_i, np.array_equal(img.cpu().numpy(), o[0].reshape(-1, 4)), np.abs(img.cpu().numpy() - o[0].reshape(-1, 4)).max(),
            np.array_equal(col.cpu().numpy(), o[1].reshape(-1, 4)), np.array_equal(pos.cpu().numpy(), o[2].reshape(-1, 4))))
    print(name, "x%d" % n, "|", " | ".join(rows), flush=True)
    ren.Cleanup()
