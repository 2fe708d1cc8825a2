"""GPU diagnostic (ONE GPU): emulate every sort-last rank of an N-way partition in turn (same shard box, same kernel
launch as that rank would issue) and print kernel ms + counters, to see what bounds a rank's launch.
    python scripts/gpu_shard_probe.py [--n 4] [--views 0,9] [--split screen]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tuvok_b200 as tb  # noqa: E402
from tuvok_b200 import _lib as L, sortlast, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--views", default="0,9")
ap.add_argument("--split", default="screen")
ap.add_argument("--only-rank", type=int, default=-1)
ap.add_argument("--repeat", type=int, default=6)
args = ap.parse_args()
w = dict(workloads.WORKLOADS["c3"])
inner = w["brick"] - 2 * w["overlap"]
finest = [-(-v // inner) for v in w["size"]]
r = tb.CudaGridLeaper(device=0, max_gpu_mem=96 << 30, hash_table_size=finest[0] * finest[1] * finest[2] * 7 + 8)
raw = torch.empty(2 * 2048 ** 3, dtype=torch.uint8, device="cuda")
r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
r.BuildVolume(raw.data_ptr(), w["brick"], w["overlap"], size=w["size"], dtype=w["dtype"], max_gradient_magnitude=0.25)
del raw
torch.cuda.empty_cache()
t1, t2 = workloads.transfer_functions(w)
r.Set1DTrans(t1); r.Set2DTrans(t2); r.SetRendermode(w["mode"]); r.SetUseLighting(True)
r.Resize(w["width"], w["height"]); r.CreateVolumePool()
fl = [np.float32(v) / np.float32(inner) for v in w["size"]]
fl = [f - f * np.finfo(np.float32).eps for f in fl]
print("view rank   ms_ray     samples    rays  brick_visits  warp_iters  max_lane_iters  Gs/s  util")
for v in [int(x) for x in args.views.split(",")]:
    r.SetRotation(workloads.orbit_rotation(v, 36))
    r._push_params()
    eye = sortlast.eye_in_volume(np.array(list(r.params.model_view)), (1.0, 1.0, 1.0))
    axes = sortlast.split_axes(0.5 - eye, args.n, args.split) if args.n > 1 else None
    boxes, _ = sortlast.shard_boxes(finest, args.n, axes)
    for g in range(args.n):
        if args.only_rank >= 0 and g != args.only_rank:
            continue
        cmin, cmax = sortlast.box_to_clip(boxes[g], finest, fl)
        r.SetShardBox(cmin, cmax)
        r.SetRotation(workloads.orbit_rotation(v, 36))
        assert r.PaintUntilConverged().converged
        r.enable_counters(True)
        r.SetRotation(workloads.orbit_rotation(v, 36))
        c = r.Paint()
        r.enable_counters(False)
        ms = []
        for _ in range(args.repeat):
            r.SetRotation(workloads.orbit_rotation(v, 36))
            ms.append(r.Paint().ms_raycast)
        m = float(np.median(ms))
        print("%4d %4d %8.3f %11d %7d %12d %11d %10d %9.2f %6.3f" %
              (v, g, m, c.samples, c.rays, c.brick_visits, c.warp_iters, c.max_lane_iters, c.samples / m / 1e6,
               c.samples / max(1.0, 32.0 * c.warp_iters)), flush=True)
