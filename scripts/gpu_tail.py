"""GPU diagnostic: is the traversal kernel throughput-bound or bound by its longest ray (critical path)?
Renders C3 views at 1, 1/2, 1/4, 1/8 resolution: a throughput-bound kernel scales with the pixel count, a
latency-bound one plateaus at (loop turns of the longest ray) x (latency of one turn).
    python scripts/gpu_tail.py [--vol 2048] [--views 0,9,18]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import tuvok_b200 as tb  # noqa: E402
from tuvok_b200 import _lib as L, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vol", type=int, default=2048)
ap.add_argument("--views", default="0,5,9,14")
ap.add_argument("--config", default="c3")
args = ap.parse_args()
w = dict(workloads.WORKLOADS[args.config])
w["size"] = (args.vol,) * 3
inner = w["brick"] - 2 * w["overlap"]
finest = [-(-v // inner) for v in w["size"]]
n_lods = 1
while max(-(-f // (1 << (n_lods - 1))) for f in finest) > 1:
    n_lods += 1
r = tb.CudaGridLeaper(device=0, max_gpu_mem=96 << 30, hash_table_size=finest[0] * finest[1] * finest[2] * n_lods + 8)
esize = {L.U8: 1, L.U16: 2, L.F32: 4}[w["dtype"]]
raw = torch.empty(args.vol ** 3 * esize, dtype=torch.uint8, device="cuda")
r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
r.BuildVolume(raw.data_ptr(), w["brick"], w["overlap"], size=w["size"], dtype=w["dtype"], max_gradient_magnitude=0.25)
del raw
t1, t2 = workloads.transfer_functions(w)
r.Set1DTrans(t1); r.Set2DTrans(t2); r.SetRendermode(w["mode"]); r.SetUseLighting(w["lighting"])
if "iso" in w:
    r.SetIsoValue(w["iso"] * {L.U8: 255.0, L.U16: 65535.0, L.F32: 1.0}[w["dtype"]])
r.Resize(w["width"], w["height"])
r.CreateVolumePool()
views = [int(v) for v in args.views.split(",")]
print("view scale   WxH        ms_ray   samples    rays  max_lane_iters warp_iters  samples/ray  Gs/s  us/turn(longest)")
for v in views:
    for scale in (1, 2, 4, 8, 16):
        W, H = w["width"] // scale, w["height"] // scale
        r.Resize(W, H)
        r.SetRotation(workloads.orbit_rotation(v, 36))
        st = r.PaintUntilConverged()
        assert st.converged
        r.enable_counters(True)
        r.SetRotation(workloads.orbit_rotation(v, 36))    # the counting kernel variant is selected when parameters are pushed
        c = r.Paint()
        r.enable_counters(False)
        r.SetRotation(workloads.orbit_rotation(v, 36))
        r.Paint()
        ms = []
        for _ in range(6):
            r.SetRotation(workloads.orbit_rotation(v, 36))    # new frame (a converged one only resumes finished rays)
            ms.append(r.Paint().ms_raycast)
        m = float(np.median(ms))
        print("%4d %5d %5dx%-5d %8.3f %10d %7d %10d %12d %10.1f %7.2f %8.3f" %
              (v, scale, W, H, m, c.samples, c.rays, c.max_lane_iters, c.warp_iters, c.samples / max(c.rays, 1),
               c.samples / m / 1e6, 1e3 * m / max(c.max_lane_iters, 1)), flush=True)
