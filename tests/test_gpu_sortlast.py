"""-m gpu: the sort-last path on one device -- every rank's partial image (CUDA renderer restricted to its brick
block with SetShardBox) against the oracle's partial image, the CUDA over operator against the oracle's, and the
binary-swap composite of the CUDA partial images against the single-GPU CUDA image (BASELINE tolerance)."""
import numpy as np
import pytest
import torch

import golden_scenes
from oracle import orc
from scene import image_diff
from tuvok_b200 import sortlast
from test_sortlast import scene_layout, run_in_process

pytestmark = pytest.mark.gpu


def cuda_partials(name, n):
    s = golden_scenes.make(name)
    finest, fl, ext = scene_layout(s)
    boxes, splits = sortlast.shard_boxes(finest, n)
    parts, oracle_parts = [], []
    for r in range(n):
        cmin, cmax = sortlast.box_to_clip(boxes[r], finest, fl)
        sr = golden_scenes.make(name, clip=(cmin, cmax))
        ren = sr.make_renderer("device")
        assert ren.PaintUntilConverged().converged
        parts.append(ren.ReadRGBA32F().reshape(-1, 4).copy())
        # only this rank's bricks were ever paged in: every resident brick touches the block
        ren.Cleanup()
        oracle_parts.append(sr.oracle_render()["image"].reshape(-1, 4).copy())
    mv, _ = s.matrices()
    eye = sortlast.eye_in_volume(mv, ext)
    n_pix = s.width * s.height
    plans = [sortlast.swap_plan(r, n, splits, boxes, finest, fl, eye, n_pix) for r in range(n)]
    return s, parts, oracle_parts, plans


@pytest.mark.parametrize("name,n", [("c2_bricked36_1d_ert", 2), ("c3_bricked36_2d_lit", 4), ("ragged_1d_lit", 8),
                                    ("c4_f32_iso", 2)])
def test_partial_images_match_oracle_and_composite_matches_single_gpu(name, n):
    s, parts, oracle_parts, plans = cuda_partials(name, n)
    for r in range(n):
        assert np.array_equal(parts[r], oracle_parts[r]), "rank %d partial image" % r
    single = s.make_renderer("device")
    assert single.PaintUntilConverged().converged
    ref8 = single.ReadRGBA8().reshape(-1, 4)

    def over(front, back):   # the CUDA compositor on device buffers
        f = torch.from_numpy(np.ascontiguousarray(front)).cuda()
        b = torch.from_numpy(np.ascontiguousarray(back)).cuda()
        o = torch.empty_like(f)
        single.composite_over(f.data_ptr(), b.data_ptr(), o.data_ptr(), f.shape[0])
        single.synchronize()
        return o.cpu().numpy()

    n_pix = s.width * s.height
    imgs = [p.copy() for p in parts]
    for k in range(len(plans[0])):
        sends = {r: imgs[r][plans[r][k]["send"][0]:plans[r][k]["send"][1]].copy() for r in range(n)}
        for r in range(n):
            rd = plans[r][k]
            lo, hi = rd["keep"]
            mine, recv = imgs[r][lo:hi], sends[rd["partner"]]
            got = over(mine, recv) if rd["i_am_front"] else over(recv, mine)
            want = orc.composite_over(mine, recv) if rd["i_am_front"] else orc.composite_over(recv, mine)
            assert np.array_equal(got, want)          # CUDA over operator == oracle's, bit for bit
            imgs[r][lo:hi] = got
    final = np.zeros((n_pix, 4), np.float32)
    for r, (a, b) in enumerate(sortlast.final_ranges(n, n_pix)):
        final[a:b] = imgs[r][a:b]
    out8 = torch.empty(n_pix * 4, dtype=torch.uint8, device="cuda")
    fin = torch.from_numpy(final).cuda()
    single.quantize_rgba8(fin.data_ptr(), out8.data_ptr(), n_pix)
    single.synchronize()
    mx, psnr = image_diff(out8.cpu().numpy().reshape(-1, 4), ref8)
    assert mx <= 2 and psnr >= 45.0, (mx, psnr)
    single.Cleanup()
