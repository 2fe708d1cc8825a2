#!/usr/bin/env python
"""Full-size parity check of a BASELINE workload (default c3): builds the workload exactly as bench.py does, then
re-traces every `--stride`-th pixel of a few orbit views with the CPU oracle on the renderer's own page table and the
pool slots the frame touched (tests/parity_gate.py).  Prints one JSON line per view and a summary line; rc != 0 when a
view is outside max |delta| <= 2/255 / PSNR >= 45 dB."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--views", default="0,12,24")
    ap.add_argument("--stride", type=int, default=8)
    ap.add_argument("--vol", type=int, default=0)
    args = ap.parse_args()
    import torch
    import tuvok_b200 as tb
    from tuvok_b200 import _lib as L, workloads
    import parity_gate

    w = dict(workloads.WORKLOADS[args.config])
    if args.vol:
        w["size"] = (args.vol,) * 3
    nx, ny, nz = w["size"]
    esize = {L.U8: 1, L.U16: 2, L.F32: 4}[w["dtype"]]
    brick = min(w["brick"], max(w["size"]) + 2 * w["overlap"])
    inner = brick - 2 * w["overlap"]
    finest = [-(-v // inner) for v in w["size"]]
    n_lods = 1
    while max(-(-f // (1 << (n_lods - 1))) for f in finest) > 1:
        n_lods += 1
    r = tb.CudaGridLeaper(device=0, max_gpu_mem=96 << 30, hash_table_size=finest[0] * finest[1] * finest[2] * n_lods + 8)
    raw = torch.empty(nx * ny * nz * esize, dtype=torch.uint8, device="cuda")
    r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
    r.BuildVolume(raw.data_ptr(), brick, w["overlap"], size=w["size"], dtype=w["dtype"], max_gradient_magnitude=0.25)
    del raw
    torch.cuda.empty_cache()
    t1, t2 = workloads.transfer_functions(w)
    r.Set1DTrans(t1); r.Set2DTrans(t2)
    r.SetRendermode(w["mode"]); r.SetUseLighting(w["lighting"])
    if "iso" in w:
        r.SetIsoValue(w["iso"] * {L.U8: 255.0, L.U16: 65535.0, L.F32: 1.0}[w["dtype"]])
    r.Resize(w["width"], w["height"])
    r.CreateVolumePool()
    bad = 0
    for v in [int(x) for x in args.views.split(",")]:
        r.SetRotation(workloads.orbit_rotation(v, 36))
        t0 = time.perf_counter()
        res = parity_gate.check_frame(r, w["size"], w["dtype"], brick, w["overlap"], stride=args.stride,
                                      threads=os.cpu_count() or 8)
        res.update(view=v, seconds=round(time.perf_counter() - t0, 2), config=args.config)
        print(json.dumps(res), flush=True)
        bad += 0 if res.get("ok") else 1
    r.Cleanup()
    print(json.dumps({"config": args.config, "views_failed": bad}), flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
