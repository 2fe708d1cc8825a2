"""Developer tool (gpurun): device time of the two quantiser passes on 2048^3-sized inputs against the HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tuvok_b200 as tb
from tuvok_b200 import _lib as L
r = tb.CudaGridLeaper()
r.set_stream(torch.cuda.current_stream().cuda_stream)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbps", 6552.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6552.0
for name, st, es, make in (("f32", L.ST_F32, 4, lambda n: torch.randn(n, device="cuda")),
                           ("i16", L.ST_I16, 2, lambda n: torch.randint(-30000, 30000, (n,), device="cuda", dtype=torch.int16)),
                           ("u32", L.ST_U32, 4, lambda n: torch.randint(0, 1 << 30, (n,), device="cuda", dtype=torch.int32))):
    n = 2048 ** 3 // (2 if es == 4 else 1)
    src = make(n)
    dst = torch.empty(n, dtype=torch.int16, device="cuda")
    best = None
    for _ in range(3):
        hist, info = r.Quantize(src.data_ptr(), st, n, 16, dst.data_ptr())
        if best is None or info.ms_range + info.ms_map < best[0] + best[1]:
            best = (info.ms_range, info.ms_map)
    print(json.dumps({"input": name, "values": n, "range_pass_ms": round(best[0], 3), "range_pass_GBps": round(n * es / 1e9 / (best[0] * 1e-3), 1),
                      "map_pass_ms": round(best[1], 3), "map_pass_GBps (read + written)": round(n * (es + 2) / 1e9 / (best[1] * 1e-3), 1),
                      "hbm_peak_GBps": peak, "map_frac": round(n * (es + 2) / 1e9 / (best[1] * 1e-3) / peak, 3)}))
    del src, dst
