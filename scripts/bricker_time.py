"""Developer tool (gpurun): time tvk_build_volume's kernels on C3's level 0 (2048^3 u16, 36^3 bricks) with the TMA box-load
path and with the generic word path (TVK_BRICKER_TMA=0), via CUDA events around the whole build (synthesis excluded)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    import tuvok_b200 as tb
    from tuvok_b200 import _lib as L, synth
    n, dt, es = int(sys.argv[2]), {"u8": L.U8, "u16": L.U16, "f32": L.F32}[sys.argv[3]], {"u8": 1, "u16": 2, "f32": 4}[sys.argv[3]]
    r = tb.CudaGridLeaper(max_gpu_mem=8 << 30)
    raw = torch.empty(n ** 3 * es, dtype=torch.uint8, device="cuda")
    r.synth_volume(raw.data_ptr(), synth.V_NOISE, (n, n, n), dt)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r.set_stream(torch.cuda.current_stream().cuda_stream)
        e0.record()
        r.BuildVolume(raw.data_ptr(), 36, 2, size=(n, n, n), dtype=dt)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    mm = r.minmax()
    import hashlib
    bricks = -(-n // 32) ** 3
    vol_b, store_b = n ** 3 * es, bricks * 36 ** 3 * es
    print(json.dumps({"mode": sys.argv[1], "n": n, "dtype": sys.argv[3], "build_ms_all_levels": round(best, 2),
                      "level0_bytes_read_plus_written_GB": round((vol_b + store_b) / 1e9, 2),
                      "minmax_sha": hashlib.sha1(mm.tobytes()).hexdigest()[:12]}))
    sys.exit(0)
for n, dt in ((2048, "u16"), (1024, "f32"), (2048, "u8")):
    for mode, env in (("tma", "1"), ("generic", "0")):
        e = dict(os.environ, TVK_BRICKER_TMA=env)
        out = subprocess.run([sys.executable, __file__, mode, str(n), dt], capture_output=True, text=True, env=e)
        print(out.stdout.strip() or out.stderr[-800:])
        print("\n".join(l for l in out.stderr.splitlines() if "level 0" in l or "level 1:" in l)[-400:])
