#!/usr/bin/env python
"""bench.py -- frames/s and Gsamples/s of the GridLeaper hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA renderer
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement (oracle/) on host cores

A "step" is one converged frame (one GLGridLeaper::Render3DRegion subframe with a fully resident
working set) of the workload `--config` (default c3: 2048^3 u16, 36^3 bricks, 2D TF + gradient
lighting, 1920x1080) from the next camera of a 36-step orbit.  N > 1 (torchrun): sort-last, the
brick blocks are sharded over the ranks and the partial images binary-swap composited
(tuvok_b200/sortlast.py); total work is fixed => "scaling": "strong".

`value`   frames/s with everything resident in HBM (device-timed with CUDA events, max over ranks)
`e2e`     frames/s through the public API with host buffers: per frame the render parameters go in
          (host) and the RGBA8 image comes back to host memory inside the timed region
`roofline` the traversal kernel's ALGORITHMIC HBM bytes / its measured duration vs the measured
          HBM peak (MEASURED_PEAKS.json); the kernel is L1/LSU-bound, see DESIGN.md
`cpu_baseline` the oracle (CPU port of the reference shader) timed on the host cores on a bounded
          sample (down-scaled volume, same TF/mode/camera), rank 0, N=1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=216)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="tvk", choices=["tvk", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c3t", "c4", "c5"])
    ap.add_argument("--no-stream", action="store_true", help="c5: skip the out-of-core orbit through the small pool")
    ap.add_argument("--path", default="gridleaper", choices=["gridleaper", "classic", "mip"],
                    help="gridleaper: GLGridLeaper page-table traversal (default); classic: per-brick GLRaycaster path; "
                         "mip: HQ MIP frame of a 2D window (GLRaycaster-MIP-Rot-FS, rotating about Y)")
    ap.add_argument("--split", default="auto", choices=["auto", "screen", "depth", "depth2", "octant", "paired", "depthw", "octantw", "pipeline"],
                    help="sort-last partition policy; auto = octant.  octant / screen run inside the library (tvk_sortlast_frame: "
                         "direct send over NCCL, octant also shards the brick store at the source); the others are the "
                         "round-1 host-driven variants (tuvok_b200/sortlast.py: binary swap, depth pipeline)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity gate (debugging)")
    ap.add_argument("--parity-stride", type=int, default=8, help="the gate re-traces every n-th pixel in x and y")
    ap.add_argument("--vol", type=int, default=0, help="override the cubic volume size (debugging)")
    ap.add_argument("--cpu-vol", type=int, default=256, help="volume size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
class CpuSample:
    """The CPU oracle (restatement of the reference GLSL, oracle/orc_render.c) on a down-scaled copy of
    the workload: same generator / TF / mode / camera, every 4th pixel of the frame in x and y."""

    def __init__(self, cfg_name, cpu_vol, threads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import orc
        self.orc = orc
        s, w, n = cpu_sample_scene(cfg_name, cpu_vol)
        self.threads = threads
        self.st = s.oracle_render(threads=threads)        # paging loop until converged (untimed)
        self.zeros = np.zeros_like(self.st["entry"])
        # the full frame has 16x the rays and (volume / cpu_vol)x the samples per ray of the sample
        self.scale = 16.0 * (w["size"][0] / float(n))
        self.kind = "port"
        self.desc = ("%d^3 down-scale of the workload volume (same generator/TF/mode/camera), %dx%d rays "
                     "(every 4th pixel of the frame), converged frame, %d threads" % (n, s.width, s.height, threads))
        # the reference's OWN shader text compiled for the host cores (oracle/_ref/glsl_baseline_<cfg>, built where the
        # reference tree is mounted by oracle/build_glsl_baseline.py): preferred over the port when it matches this scene
        self.glsl = None
        exe = os.path.join(ROOT, "oracle", "_ref", "glsl_baseline_%s" % cfg_name)
        sig = baseline_signature(s, self.st, cfg_name, n)
        try:
            if os.path.exists(exe) and json.load(open(exe + ".json")) == sig:
                import glsl_ref
                import tempfile
                self._tmp = tempfile.mkdtemp(prefix="tvk_glsl_")
                p = self.st["params"]
                u = orc.uniforms(p)
                glsl_ref.scene_file(os.path.join(self._tmp, "scene.bin"), p, u, orc.ray_exit_eye(p), self.st["entry"], self.zeros,
                                    self.st["covered"], self.st["meta"], self.st["pool"].meta_dim, self.st["atlas"], self.st["tf"])
                self.glsl = exe
                self.kind = "reference"
                _, rs = orc.raycast(p, self.st["atlas"], self.st["meta"], self.st["tf"], self.st["entry"], self.zeros,
                                    self.st["exit"], self.st["covered"], None, threads)
                self.samples = int(rs.samples)            # identical sample sequence (tests/test_glsl_ref.py)
                self.desc = ("the reference's GLGridLeaper GLSL (blend + Method + GradientTools + lighting + Compositing + the "
                             "GLSL generated by GLVolumePool/GLHashTable) compiled for the host cores (OpenMP over fragments, "
                             "oracle/glsl/glsl_emu.h) on a " + self.desc)
        except Exception as e:      # fall back to the port, say why
            self.desc += " [executed-GLSL baseline unavailable: %s]" % e

    def frame(self):
        """one converged pass on the host cores; returns (seconds, samples)"""
        st = self.st
        if self.glsl:
            env = dict(os.environ, OMP_NUM_THREADS=str(self.threads))
            out = subprocess.run([self.glsl, os.path.join(self._tmp, "scene.bin"), os.path.join(self._tmp, "out.bin"), "1"],
                                 capture_output=True, text=True, check=True, env=env).stdout.split()
            return float(out[0]), self.samples
        return self.frame_port()

    def frame_port(self):
        """the same pass through the oracle port (C restatement, OpenMP); returns (seconds, samples)"""
        st = self.st
        t0 = time.perf_counter()
        _, rs = self.orc.raycast(st["params"], st["atlas"], st["meta"], st["tf"], st["entry"], self.zeros,
                                 st["exit"], st["covered"], None, self.threads)
        return time.perf_counter() - t0, int(rs.samples)


def cpu_sample_scene(cfg_name, cpu_vol):
    """The bounded CPU sample of a workload: down-scaled volume, every 4th pixel, same generator / TF / mode / camera."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from scene import Scene
    from tuvok_b200 import workloads
    w = workloads.WORKLOADS[cfg_name]
    n = min(cpu_vol, w["size"][0])
    t1, t2 = workloads.transfer_functions(w)
    iso = w.get("iso", 0.5) * {0: 255.0, 1: 65535.0, 2: 1.0}[w["dtype"]]
    sw, sh = max(64, w["width"] // 4), max(64, w["height"] // 4)
    s = Scene(kind=w["kind"], size=(n, n, n), dtype=w["dtype"], brick=min(w["brick"], n + 4),
              overlap=w["overlap"], mode=w["mode"], lighting=w["lighting"], width=sw, height=sh,
              rotation=workloads.orbit_rotation(3), tf2d=t2, isovalue=iso, max_gpu_mem=8 << 30)
    s.tf1d = t1
    return s, w, n


def baseline_signature(s, st, cfg_name, n):
    """What the scene-specific generated GLSL depends on; stored next to the prebuilt executable."""
    pool = st["pool"]
    p = st["params"]
    return {"config": cfg_name, "volume": n, "brick": int(s.brick[0]), "overlap": int(s.overlap), "dtype": int(s.dtype),
            "pool_size": [int(v) for v in pool.pool_size], "meta_dim": [int(v) for v in pool.meta_dim],
            "mode": int(s.mode), "lighting": bool(s.lighting), "hash_size": int(p.hash_size), "rehash": int(p.rehash_count),
            "strategy": int(s.strategy), "image": [int(s.width), int(s.height)]}


def run_reference(args, rank):
    """--impl reference: the reference's algorithm on the host cores (the reference GL renderer cannot
    be built or run: no GL/Qt/bison here or on the GPU box) == the oracle port, all threads."""
    if rank != 0:
        return
    from tuvok_b200 import workloads
    threads = os.cpu_count() or 1
    cs = CpuSample(args.config, args.cpu_vol, threads)
    t_begin = time.perf_counter()
    for _ in range(args.warmup):
        cs.frame()
        if time.perf_counter() - t_begin > 60:
            break
    times, samples = [], 0
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        dt, samples = cs.frame()
        times.append(dt)
        if time.perf_counter() - t_begin > 150:
            break
    frame_s = float(np.mean(times))
    sps = samples / frame_s
    # frames/s of the FULL workload at the measured CPU sample rate: samples/s / samples of one full frame (a
    # device-counted constant of the seeded workload, tuvok_b200/workloads.py) -- the same conversion as the
    # cpu_baseline object of the GPU arm.  Fallback for workloads without a recorded count: rays x depth scaling.
    spf = workloads.SAMPLES_PER_FRAME.get(args.config) if not args.vol else None
    fps_equiv = sps / spf if spf else 1.0 / (frame_s * cs.scale)
    how = ("value = CPU samples/s / %.4g samples of one full frame" % spf) if spf else \
          ("value = 1 / (step time x %.0f)" % cs.scale)
    w = workloads.WORKLOADS[args.config]
    port = {}
    if cs.kind == "reference":      # the restated port on the same sample, reported beside the executed GLSL
        pt, ps = min(cs.frame_port() for _ in range(2))
        port = {"port_gsamples_per_s": ps / pt / 1e9,
                "port_value": (ps / pt / spf) if spf else 1.0 / (pt * cs.scale)}
    line = {"impl": "reference", "metric": "frames_per_s", "value": fps_equiv, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": frame_s * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "gsamples_per_s": sps / 1e9,
            "config": {"workload": w["label"], "bounded_sample": cs.desc,
                       "note": "each step = one bounded-sample frame; " + how + ", the frame rate of the "
                               "full-resolution, full-size workload at the measured CPU sample rate"},
            "cpu_baseline": {"value": fps_equiv, "unit": "frames/s", "cores": threads, "kind": cs.kind,
                             "sample": cs.desc, "gsamples_per_s": sps / 1e9, **port},
            "e2e": {"value": fps_equiv, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
PARITY_VIEWS = (0, 12, 24)      # orbit views re-traced by the oracle (and compared with the single-GPU frame at N > 1)


def workload_geometry(w):
    brick = min(w["brick"], max(w["size"]) + 2 * w["overlap"])
    inner = brick - 2 * w["overlap"]
    finest = [-(-v // inner) for v in w["size"]]
    n_lods = 1
    while max(-(-f // (1 << (n_lods - 1))) for f in finest) > 1:
        n_lods += 1
    flayout = [np.float32(v) / np.float32(inner) for v in w["size"]]
    flayout = [f - f * np.finfo(np.float32).eps if float(int(f)) == float(f) else f for f in flayout]
    ext = np.array(w["size"], np.float64)
    return brick, inner, finest, n_lods, flayout, ext / ext.max()


def host_mem_available():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


PROC_INFO = {}


def procedural_minmax_table(r, w, brick, rank, world, dist):
    """min/max table of a procedural workload: every rank evaluates a slice on its device, the slices are gathered"""
    import ctypes
    import torch
    from tuvok_b200 import _lib as L
    n, lods = ctypes.c_uint64(), ctypes.c_uint32()
    L.lib().tvk_procedural_brick_count(L.u32x3(*w["size"]), L.u32x3(brick, brick, brick), w["overlap"], ctypes.byref(n), ctypes.byref(lods))
    n = int(n.value)
    per = -(-n // world)
    first, count = min(n, rank * per), max(0, min(n, (rank + 1) * per) - min(n, rank * per))
    t0 = time.perf_counter()
    part = r.procedural_minmax(w["kind"], w["size"], w["dtype"], brick, w["overlap"], first, count)
    dt = time.perf_counter() - t0
    if world > 1:
        pad = np.zeros((per, 4), np.float64)
        pad[:count] = part
        mine = torch.from_numpy(pad).cuda()
        allp = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        part = torch.cat(allp)[:n].cpu().numpy()
    PROC_INFO.update(bricks=n, lods=int(lods.value), minmax_pass_s=round(dt, 2), minmax_bricks_per_rank=count)
    return part


def make_renderer(w, device, stream, shard=None, rank=0, world=1, dist=None, minmax=None):
    """The workload on one renderer: synthetic volume bricked on the device (tvk_build_volume), transfer functions, mode,
    pool.  shard = (clip_min, clip_max): only the bricks of that block are kept in the brick store (sort-last at the
    source, tvk_set_store_shard)."""
    import torch
    import tuvok_b200 as tb
    from tuvok_b200 import _lib as L
    from tuvok_b200 import workloads
    brick, inner, finest, n_lods, _, _ = workload_geometry(w)
    nx, ny, nz = w["size"]
    esize = {L.U8: 1, L.U16: 2, L.F32: 4}[w["dtype"]]
    hash_size = finest[0] * finest[1] * finest[2] * n_lods + 8   # collision-free: one slot per serialised id
    # the pool budget counts voxels as the reference does; 8 / 16-bit pools take twice that in HBM (x-pair layout)
    r = tb.CudaGridLeaper(device=device, max_gpu_mem=(40 << 30) if w.get("procedural") else (96 << 30), hash_table_size=hash_size)
    r.set_stream(stream.cuda_stream)
    if w.get("procedural"):
        # the dataset is never materialised: bricks come from the host generator threads through the pinned streaming path
        if minmax is None:
            minmax = procedural_minmax_table(r, w, brick, rank, world, dist)
        local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        cache = int(min(32 << 30, 0.4 * host_mem_available() / local))
        threads = max(2, (os.cpu_count() or 8) // local)
        r.SetProceduralVolume(w["kind"], w["size"], w["dtype"], brick, w["overlap"], minmax=minmax, host_cache_bytes=cache,
                              threads=threads, max_gradient_magnitude=0.25)
        PROC_INFO.update(host_cache_GiB=round(cache / 2 ** 30, 1), generator_threads=threads)
        r._minmax_table = minmax
    else:
        if shard is not None:
            r.SetStoreShard(*shard)
        raw = torch.empty(nx * ny * nz * esize, dtype=torch.uint8, device="cuda")
        r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
        r.BuildVolume(raw.data_ptr(), brick, w["overlap"], size=w["size"], dtype=w["dtype"], max_gradient_magnitude=0.25)
        del raw
        torch.cuda.empty_cache()
    t1, t2 = workloads.transfer_functions(w)
    r.Set1DTrans(t1)
    r.Set2DTrans(t2)
    r.SetRendermode(w["mode"])
    r.SetUseLighting(w["lighting"])
    if "iso" in w:
        r.SetIsoValue(w["iso"] * {L.U8: 255.0, L.U16: 65535.0, L.F32: 1.0}[w["dtype"]])
    r.Resize(w["width"], w["height"])
    if "translate" in w:
        r.SetTranslation(tb.translation(*w["translate"]))
    r.CreateVolumePool()
    return r


def merge_parity(results):
    """Worst case over the checked views (and, through an all-reduce by the caller, over the ranks)."""
    ok = all(x.get("ok") for x in results)
    psnrs = [x["psnr_db"] for x in results if "psnr_db" in x]
    fin = [p for p in psnrs if p != "inf"]
    return {"ok": bool(ok), "max_abs_255": max([x.get("max_abs_255", 255) for x in results] or [255]),
            "psnr_db": (min(fin) if fin else "inf"), "pixels": int(sum(x.get("pixels", 0) for x in results)),
            "float_bit_identical": all(x.get("float_bit_identical", False) for x in results),
            "views": len(results), "stride": results[0].get("stride") if results else None,
            "checker": "CPU oracle (oracle/orc_render.c) re-tracing every stride-th pixel on the renderer's own page table "
                       "and the touched pool slots read back from the device (tests/parity_gate.py); gate: max <= 2/255, PSNR >= 45 dB",
            "errors": [x["error"] for x in results if "error" in x]}


# Sort-last composite against the single-GPU frame.  The chain that is checked BIT FOR BIT is: every rank's partial image ==
# the oracle's partial image (`parity`, scope "every rank's partial image") and the n-way over kernel == orc_over
# (tests/test_gpu_sortlast.py).  The composite itself cannot equal the single-GPU frame bit for bit, for two reasons the
# reference shares (DESIGN.md sections 4, 5): (1) sample positions live in pool coordinates, i.e. relative to the slot a
# brick happens to occupy, and a rank's pool holds other bricks in other slots than the single GPU's -- positions move
# by an ulp and now and then a sample falls on the other side of a transfer-function edge (nearest-neighbour tables):
# that pixel changes by up to ONE sample's contribution; (2) a back rank accumulates its block without knowing the
# alpha in front of it, so where the single-GPU ray ends inside a back block the over operator can only cut the back
# image uniformly (<= 0.0086 = 2.2/255 before rounding).  The gate is therefore: PSNR >= 45 dB over all pixels,
# max |delta| <= 2/255 on all pixels but an outlier budget of 1e-5 of them, and no pixel further off than one sample's
# largest contribution (the table's maximal opacity) + 2; `ok_strict` says whether max <= 2/255 held on every pixel.
COMPOSITE_OUTLIER_FRACTION = 1e-5
# C5 (camera inside the volume, translucent table, rays of ~1300 samples that saturate only after crossing up to four
# blocks): the uniform cut of mechanism (2) applies at every face a saturating ray crosses, the errors add up
# (measured at N = 8: 4.3e-4 of the pixels beyond 2/255, max 8/255, PSNR 55.5 dB) -- the budget is stated per workload
COMPOSITE_OUTLIER_FRACTION_BY_WORKLOAD = {"c5": 1e-3}


def composite_gate(res, world, hard_max, fraction=COMPOSITE_OUTLIER_FRACTION):
    import math
    import parity_gate
    n = sum(x["pixels"] for x in res)
    mx = max(x["max_abs_255"] for x in res)
    psnr = min(x["psnr"] for x in res)
    over2 = sum(x["over_2"] for x in res)
    budget = int(math.ceil(fraction * n))
    ok = psnr >= parity_gate.MIN_PSNR_DB and over2 <= budget and mx <= hard_max
    return {"ok": bool(ok), "ok_strict": bool(ok and mx <= parity_gate.MAX_ABS_255), "max_abs_255": int(mx),
            "psnr_db": "inf" if math.isinf(psnr) else round(psnr, 2), "pixels": int(n), "views": len(res),
            "pixels_over_1": int(sum(x["over_1"] for x in res)), "pixels_over_2": int(over2), "outlier_budget": budget,
            "worst": [e for x in res for e in x["worst"]][:8],
            "checker": ("the gathered %d-rank frame (RGBA8, all pixels) against the frame of the same view rendered by ONE GPU "
                        "through the same library before the store was sharded; gate: PSNR >= 45 dB, max <= 2/255 on all but "
                        "%g of the pixels (slot-dependent sample positions at transfer-function edges, early ray termination across "
                        "block faces: DESIGN.md section 5), none above one sample's largest contribution + 2 = %d/255"
                        % (world, fraction, hard_max))}


def run_tvk(args, rank, world, local_rank):
    import torch
    from tuvok_b200 import _lib as L
    from tuvok_b200 import sortlast, workloads
    sys.path.insert(0, os.path.join(ROOT, "tests"))

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    w = dict(workloads.WORKLOADS[args.config])
    if os.environ.get("TVK_TZ"):           # developer overrides for camera-path probes
        w["translate"] = (0.0, 0.0, float(os.environ["TVK_TZ"]))
    if os.environ.get("TVK_ALPHA_SCALE"):
        w["alpha_scale"] = float(os.environ["TVK_ALPHA_SCALE"])
    if args.vol:
        w["size"] = (args.vol,) * 3
        w["label"] = w["label"].replace(str(workloads.WORKLOADS[args.config]["size"][0]) + "^3", "%d^3" % args.vol)
    brick, inner, finest, n_lods, flayout, ext = workload_geometry(w)
    esize = {L.U8: 1, L.U16: 2, L.F32: 4}[w["dtype"]]
    stream = torch.cuda.current_stream()
    n_pixels = w["width"] * w["height"]
    n_views = 36
    mip = args.path == "mip"
    classic = args.path in ("classic", "mip")          # the per-brick paths: one converged frame per call
    if classic and world > 1:
        raise SystemExit("the classic / MIP paths are single-GPU (sort-last shards the GridLeaper path)")
    split = args.split if args.split != "auto" else "octant"
    legacy = world > 1 and split not in ("octant", "screen", "paired")     # round-1 host-driven policies (binary swap / pipeline)
    do_parity = not args.no_parity and not classic

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_min(flag):
        if dist is None:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item()) > 0.5

    t_setup = time.perf_counter()
    proc_table = None
    if w.get("procedural"):      # min/max table first (collective: every rank evaluates a slice on its device)
        import tuvok_b200 as tb
        tmp = tb.CudaGridLeaper(device=local_rank)
        proc_table = procedural_minmax_table(tmp, w, brick, rank, world, dist)
        tmp.Cleanup()
        del tmp
    # ---- N > 1: the single-GPU frames the composite is compared with (rank 0, whole store, before sharding) --------
    ref_frames = {}
    if world > 1 and do_parity and rank == 0:
        r0 = make_renderer(w, local_rank, stream, minmax=proc_table)
        for v in PARITY_VIEWS:
            r0.SetRotation(workloads.orbit_rotation(v, n_views))
            if not r0.PaintUntilConverged().converged:
                raise RuntimeError("single-GPU reference view %d did not converge" % v)
            r0._dirty = True      # a fresh pass over the now resident pool, as the timed frames are: the floats of a frame that
            r0.Paint()            # resumed rays after paging depend on the paging history
            ref_frames[v] = r0.ReadRGBA8().copy()
        r0.Cleanup()
        del r0
        torch.cuda.empty_cache()
    barrier()

    # ---- this rank's renderer -----------------------------------------------------------------------------------
    shard = None
    lib_sl = world > 1 and not legacy
    if lib_sl and split == "octant" and not w.get("procedural"):
        # view-independent blocks: the brick store is sharded at the source (memory per GPU falls with N)
        cmin, cmax, _ = sortlast.plan(finest, flayout, ext, np.eye(4, dtype=np.float32) + 0, world, L.SL_OCTANT)
        shard = (tuple(float(v) for v in cmin[rank]), tuple(float(v) for v in cmax[rank]))
    r = make_renderer(w, local_rank, stream, shard, rank, world, dist, minmax=proc_table)
    info = r.info()
    pipe = sl = None
    if lib_sl:
        sortlast.init_library_sortlast(r, rank, world, dist, {"octant": L.SL_OCTANT, "screen": L.SL_SCREEN, "paired": L.SL_PAIRED}[split])
    elif legacy and split == "pipeline":
        pipe = sortlast.DepthPipeline(r, rank, world, finest, flayout, ext, align=1)
        r.SetRotation(workloads.orbit_rotation(0, 36))
        r.Paint()
        pipe.set_weights()
    elif legacy:
        sl = sortlast.SortLastRenderer(r, rank, world, finest, flayout, ext, policy=split)
        if split.endswith("w"):
            r.SetRotation(workloads.orbit_rotation(0, 36))
            r.Paint()
            sl.set_weights()

    def paint_per_brick():
        return r.PaintHQMIP("coronal") if mip else r.PaintClassic()

    def set_view(i):
        r.SetRotation(workloads.orbit_rotation(i % n_views, n_views))
        if mip:
            r.SetMIPRotationAngle(10.0 * (i % n_views))    # GLRenderer::SetMIPRotationAngle: the 2D window's MIP turntable
        if pipe is not None:
            pipe.view_id = i % n_views
        if sl is not None:
            sl.update_partition()

    sl_stats = []
    sl_mode = [0]
    lpt_on = 0 if os.environ.get("TVK_TILE_LPT", "1").startswith("0") else 1

    def frame(i):
        """one step on this rank; returns the stats of the (single) subframe"""
        set_view(i)
        if lib_sl:
            st = r.SortLastFrame()
            sl_stats.append((st.frame.ms_raycast, st.ms_exchange, st.ms_frame, st.ms_wait_peers))
            sl_mode[0] = int(st.peer_memory)
            return st.frame
        if pipe is not None:
            return pipe.render_frame()[0]
        if sl is None:
            return paint_per_brick() if classic else r.Paint()
        lo, hi, img, st = sl.render()
        sl.gather(lo, hi, img)
        return st

    # ---- out-of-core workloads: cold start = first frame to converged on an empty pool and an empty host cache ---------
    cold = None
    if w.get("procedural"):
        barrier()
        tc = time.perf_counter()
        n_sub, n_paged = 0, 0
        for _ in range(256):
            st = frame(0)
            n_sub += 1; n_paged += st.bricks_paged
            if all_min(st.converged):
                break
        barrier()
        ss = r.stream_stats()
        cold = {"first_frame_to_converged_s": round(time.perf_counter() - tc, 3), "subframes": n_sub, "bricks_paged_this_rank": int(n_paged),
                "bricks_generated_this_rank": int(ss.bricks_generated),
                "generator_ms_per_brick_and_thread": round(ss.source_thread_ms / max(1, ss.bricks_generated), 2)}

    # ---- setup (untimed): page the working set of every orbit view in --------------------------
    paged = 0
    for i in range(n_views):
        if world > 1:
            for _ in range(64):                     # collective subframes until every rank has converged
                st = frame(i)
                paged += st.bricks_paged
                if all_min(st.converged):
                    break
            else:
                raise RuntimeError("view %d did not converge on every rank" % i)
            continue
        set_view(i)
        st = paint_per_brick() if classic else r.PaintUntilConverged()
        paged += st.bricks_paged
        if not st.converged:
            raise RuntimeError("view %d did not converge (pool too small?)" % i)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    # ---- parity gate (untimed): this rank's frames re-traced by the oracle ---------------------------------------
    parity = None
    if do_parity:
        import parity_gate
        res = []
        for v in PARITY_VIEWS:
            set_view(v)
            for which in range(r._sl_blocks_per_rank if lib_sl else 1):
                if lib_sl:
                    bmin, bmax = r.SortLastBlockOf(which)    # the rank's block(s) for this view, as the library cuts them
                    r.SetShardBox(bmin, bmax)
                res.append(parity_gate.check_frame(r, w["size"], w["dtype"], brick, w["overlap"], stride=args.parity_stride,
                                                   threads=max(1, (os.cpu_count() or 8) // max(1, world))))
        if lib_sl:
            r.SetShardBox((0, 0, 0), (1, 1, 1))
        parity = merge_parity(res)
        if dist is not None:      # worst case over the ranks
            t = torch.tensor([0.0 if parity["ok"] else 1.0, float(parity["max_abs_255"]),
                              -(1e9 if parity["psnr_db"] == "inf" else float(parity["psnr_db"])),
                              0.0 if parity["float_bit_identical"] else 1.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            px = torch.tensor([float(parity["pixels"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(px, op=dist.ReduceOp.SUM)
            bad, mx, npsnr, nbit = (float(v) for v in t.cpu())
            parity.update(ok=bad == 0.0, max_abs_255=int(mx), psnr_db=("inf" if -npsnr >= 1e9 else round(-npsnr, 2)),
                          float_bit_identical=nbit == 0.0, pixels=int(px.item()),
                          scope="every rank's partial image (rays restricted to its brick block) against the oracle's partial image")

    # ---- counting pass (untimed): samples / bricks touched per view -----------------------------
    r.enable_counters(True)
    samples, rays, touched, visits = [], [], [], []
    alive_it, warp_it = 0, 0
    for i in range(n_views):
        st = frame(i)
        samples.append(st.samples); rays.append(st.rays); touched.append(st.bricks_touched); visits.append(st.brick_visits)
        alive_it += st.alive_lane_iters; warp_it += st.warp_iters
    r.enable_counters(False)

    # ---- fetch-path ceiling of the traversal kernel on this pool (untimed) ------------------------
    fetch_peak = None
    if not classic:
        try:
            set_view(0)
            rates = [r.probe_fetch(256, d)[0] for d in ((0.93, 0.3, 0.2), (0.5, 0.62, 0.6), (0.2, 0.3, 0.93))]
            fetch_peak = float(np.mean(rates))
        except Exception:
            fetch_peak = None

    # ---- value: device-timed resident frames ----------------------------------------------------
    for i in range(args.warmup):
        frame(i)
    sl_stats.clear()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms_ray, not_conv = 0.0, 0
    e0.record(stream)
    for i in range(args.steps):
        st = frame(args.warmup + i)
        ms_ray += st.ms_raycast
        not_conv += 0 if st.converged else 1
    if lib_sl:
        r.SortLastFlush()      # overlapped exchange: the end event waits for the last frame's blend and gather as well
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if not_conv and pipe is None:
        raise RuntimeError("%d timed frames were not converged" % not_conv)
    timed_sl = list(sl_stats)

    # ---- e2e: public API with host buffers (params in, RGBA8 image out to host) ----------------
    pinned = [r.host_alloc((w["height"], w["width"], 4)) for _ in range(2)] if sl is None else None
    last_stage = pipe is not None and rank == world - 1
    pending = []
    checksum = 0
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        set_view(args.warmup + i)
        if lib_sl:
            # the gathered frame leaves rank 0: PBO-style double-buffered read-back there, every frame
            r.SortLastFrame()
            if rank == 0:
                r.SortLastReadRGBA8Async(pinned[i % 2])
                r.WaitRead(pending_allowed=1)
                if i > 0:
                    checksum += int(pinned[(i - 1) % 2][w["height"] // 2, w["width"] // 2, 3])
        elif pipe is not None:
            pipe.render_frame()
            if last_stage:
                r.ReadRGBA8Async(pinned[i % 2])
                r.WaitRead(pending_allowed=1)
                if i > 0:
                    checksum += int(pinned[(i - 1) % 2][w["height"] // 2, w["width"] // 2, 3])
        elif sl is None:
            # PBO-style double-buffered read-back: frame i is copied to pinned host memory while frame i+1 renders;
            # every frame's RGBA8 image is in host memory (and touched) before the timed region ends
            if classic:
                paint_per_brick()
            else:
                r.Paint()
            r.ReadRGBA8Async(pinned[i % 2])
            r.WaitRead(pending_allowed=1)
            if i > 0:
                checksum += int(pinned[(i - 1) % 2][w["height"] // 2, w["width"] // 2, 3])
        else:
            lo, hi, img, _ = sl.render()
            full = sl.gather(lo, hi, img)
            if rank == 0:
                pending.append(sl.read_rgba8_async(full))
                if len(pending) > 1:
                    img_h, ev = pending.pop(0)
                    ev.synchronize()
                    checksum += int(img_h[(w["height"] // 2) * w["width"] + w["width"] // 2, 3])
    for img_h, ev in pending:
        ev.synchronize()
        checksum += int(img_h[(w["height"] // 2) * w["width"] + w["width"] // 2, 3])
    if sl is None and (pipe is None or last_stage) and (not lib_sl or rank == 0):
        r.WaitRead(0)
        checksum += int(pinned[(args.steps - 1) % 2][w["height"] // 2, w["width"] // 2, 3])
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- out-of-core workloads: what the streaming path moved, and (N = 1) an orbit through a pool that holds one view's
    # working set but not the orbit's, so that every view change pages bricks in from host memory ---------------------
    stream_obj = None
    if w.get("procedural"):
        ss = r.stream_stats()

        def snap(st):
            return {"bricks_uploaded": int(st.bricks_uploaded), "h2d_GB": round(st.h2d_bytes / 1e9, 3),
                    "h2d_device_ms": round(st.h2d_ms, 2), "upload_wall_ms": round(st.upload_ms, 1),
                    "h2d_GBps_while_copying": round(st.h2d_bytes / 1e9 / max(1e-9, st.h2d_ms * 1e-3), 2),
                    "bricks_generated": int(st.bricks_generated), "host_cache_hits": int(st.host_cache_hits),
                    "generator_thread_ms": round(st.source_thread_ms, 1), "generator_threads": int(st.source_threads),
                    "host_cache_page_locked": bool(st.host_cache_pinned)}
        stream_obj = {"setup_and_cold_start_this_rank": snap(ss)}
        # pinned H2D peak of this box, measured the same way (1 GiB pinned -> device, events on the stream)
        hp = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
        dp = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        dp.copy_(hp, non_blocking=True); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(4):
            dp.copy_(hp, non_blocking=True)
        a1.record(); torch.cuda.synchronize()
        stream_obj["pinned_h2d_peak_GBps"] = round(4 * (1 << 30) / 1e9 / (a0.elapsed_time(a1) * 1e-3), 2)
        del hp, dp
        if world == 1 and not args.no_stream:
            slots = int(max(touched) * 1.5) + 64
            side = int(np.ceil(slots ** (1.0 / 3.0)))
            r.CreateVolumePool((brick * side, brick * side, brick * side))
            for i in range(n_views):              # first lap: the host cache already holds the orbit's bricks (setup)
                set_view(i)
                r.PaintUntilConverged()
            s0 = r.stream_stats()
            torch.cuda.synchronize()
            tq = time.perf_counter()
            sub = 0
            for i in range(n_views):
                set_view(i)
                st = r.PaintUntilConverged()
                if not st.converged:
                    raise RuntimeError("out-of-core lap: view %d did not converge in a pool of %d slots" % (i, side ** 3))
            torch.cuda.synchronize()
            lap = time.perf_counter() - tq
            s1 = r.stream_stats()
            nb = int(s1.bricks_uploaded - s0.bricks_uploaded)
            by = s1.h2d_bytes - s0.h2d_bytes
            stream_obj["out_of_core_orbit"] = {
                "pool_slots": side ** 3, "pool_GB": round(side ** 3 * brick ** 3 * esize * (2 if w["dtype"] != L.F32 else 1) / 1e9, 2),
                "largest_view_working_set_bricks": int(max(touched)), "frames_per_s": round(n_views / lap, 2),
                "bricks_paged_per_view": round(nb / n_views, 1), "h2d_GB_per_view": round(by / 1e9 / n_views, 3),
                "h2d_GBps_while_copying": round(by / 1e9 / max(1e-9, (s1.h2d_ms - s0.h2d_ms) * 1e-3), 2),
                "h2d_GBps_over_upload_wall_time": round(by / 1e9 / max(1e-9, (s1.upload_ms - s0.upload_ms) * 1e-3), 2),
                "host_cache_hit_fraction": round((s1.host_cache_hits - s0.host_cache_hits) / max(1, nb), 3),
                "note": "converged frames (all subframes) of a 36-view orbit through a pool 1.5x the largest single-view working "
                        "set; every view change pages its bricks in from the host brick cache through pinned staging"}

    # ---- N > 1: the gathered composite against the single-GPU frame of the same view (untimed) --------------------
    composite = None
    if lib_sl and do_parity:
        import parity_gate
        res = []
        for v in PARITY_VIEWS:
            for _ in range(64):
                st = frame(v)
                if all_min(st.converged):
                    break
            # a frame must never be read one exchange late (overlapped exchange): render ANOTHER view, then exactly one frame
            # of this one, and read that
            for _ in range(64):
                st = frame(v + 6)
                if all_min(st.converged):
                    break
            st = frame(v)
            if not all_min(st.converged):
                raise RuntimeError("view %d pages bricks again after it had converged" % v)
            if rank == 0:
                got = r.SortLastReadRGBA8()
                ref = ref_frames[v]
                mx, psnr = parity_gate.image_metrics(got, ref)
                d8 = np.abs(got.astype(np.int32) - ref.astype(np.int32)).reshape(w["height"], w["width"], 4).max(axis=-1)
                worst = [{"view": int(v), "x": int(x), "y": int(y), "composite": [int(c) for c in got.reshape(w["height"], w["width"], 4)[y, x]],
                          "single": [int(c) for c in ref.reshape(w["height"], w["width"], 4)[y, x]]}
                         for y, x in np.argwhere(d8 > parity_gate.MAX_ABS_255)[:4]]
                res.append({"max_abs_255": mx, "psnr": psnr, "pixels": int(d8.size), "over_1": int((d8 > 1).sum()),
                            "over_2": int((d8 > parity_gate.MAX_ABS_255).sum()), "worst": worst})
        if rank == 0:
            composite = composite_gate(res, world, int(w.get("alpha_max", 16)) + 2,
                                       COMPOSITE_OUTLIER_FRACTION_BY_WORKLOAD.get(args.config, COMPOSITE_OUTLIER_FRACTION))

    times = torch.tensor([ms_total, ms_ray, e2e_s * 1e3, float(not_conv)], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(np.sum(samples)), float(np.sum(touched))], dtype=torch.float64, device="cuda")
    per_rank = None
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        if timed_sl:
            mine = torch.tensor([float(np.mean([x[0] for x in timed_sl])), float(np.mean([x[1] for x in timed_sl])),
                                 float(np.mean([x[2] for x in timed_sl])), float(np.mean([x[3] for x in timed_sl])),
                                 float(np.sum(samples)) / n_views,
                                 float(torch.cuda.max_memory_allocated() / 2 ** 30)], dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = [[round(float(v), 4) for v in t.cpu()] for t in allr]
    ms_total, ms_ray, e2e_ms, not_conv_max = (float(v) for v in times.cpu())
    samples_per_orbit, touched_per_orbit = (float(v) for v in tot.cpu())

    rc = 0
    if rank == 0:
        k = args.steps
        fps = k / (ms_total * 1e-3)
        # samples of the timed steps: views cycle through the orbit
        view_ids = [(args.warmup + i) % n_views for i in range(k)]
        if world == 1:
            step_samples = float(sum(samples[v] for v in view_ids))
            step_touched = float(sum(touched[v] for v in view_ids))
        else:
            step_samples = samples_per_orbit * k / n_views
            step_touched = touched_per_orbit * k / n_views
        gsps = step_samples / (ms_total * 1e-3) / 1e9
        slot_bytes = brick ** 3 * esize
        # algorithmic HBM bytes per launch: every sampled brick once + page-table entries of the
        # visited bricks + the kernel's per-pixel outputs (acc colour, resume colour, resume position)
        out_bytes = 48.0 if not classic else (24.0 if mip else 16.0)   # per pixel: 3 MRTs / colour (+ MIP maximum image)
        alg_bytes = (step_touched * slot_bytes + step_touched * 4 + k * n_pixels * out_bytes) / k / world
        ray_ms = ms_ray / k
        peak, which = peaks()
        achieved = alg_bytes / (ray_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "raycast_traffic.json")
        if os.path.exists(tp) and not classic:
            try:
                traffic = json.load(open(tp)).get(args.config)
            except Exception:
                traffic = None
        # fetch path: filter taps per sample x 8 voxels x bytes.  DVR with a 2D table or lighting: 7 taps;
        # 1D unlit and the isosurface march: 1 tap (the 7-tap gradient is taken once per ray, at the hit)
        taps = 7 if (w["mode"] == 1 or (w["mode"] == 0 and w["lighting"])) and not mip else 1
        kern_sps = step_samples / k / world / (ray_ms * 1e-3)            # per GPU, inside the traversal kernel
        fetch = {"taps_per_sample": taps, "bytes_per_sample": taps * 8 * esize,
                 "achieved_gsamples_per_s": kern_sps / 1e9, "achieved_gbs": kern_sps * taps * 8 * esize / 1e9,
                 "peak_gsamples_per_s": (fetch_peak / 1e9) if fetch_peak else None,
                 "peak_gbs": (fetch_peak * taps * 8 * esize / 1e9) if fetch_peak else None,
                 "frac": (kern_sps / fetch_peak) if fetch_peak else None,
                 "peak_source": "measured in this run: tvk_probe_fetch = the kernel's own footprint loads + filter trees on the "
                                "resident pool, nothing else in the loop (3 march directions, mean)"}
        par = "single GPU"
        if lib_sl:
            how = ("partial images read straight out of the peers' memory over NVLink by the n-way over kernel, RGBA8 slices stored "
                   "into rank 0's frame by the same kernel, flag words instead of collectives (no NCCL call on the frame's path)"
                   if sl_mode[0] else "direct-send RGBA32F slices in one NCCL group, n-way over kernel, RGBA8 gather on rank 0")
            if sl_mode[0] == 2:
                how += ("; the exchange of frame f runs on a second stream while frame f + 1 is traversed (a rank runs up to one "
                        "frame ahead; the timed region ends after the last frame's gather)")
            par = ("sort-last x%d inside the library (tvk_sortlast_frame): %s brick blocks%s, %s" %
                   (world, split, ", brick store sharded at the source" if shard is not None else "", how))
        elif pipe is not None:
            par = ("depth pipeline x%d (rank s = slab s from the eye, hand-over of resume position + colour over NCCL, "
                   "slabs balanced by non-empty bricks; frames in flight = %d)" % (world, world))
        elif world > 1:
            par = "sort-last x%d (binary swap driven from the host, %s partition)" % (world, split)
        line = {
            "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": k,
            "warmup": args.warmup, "ms_per_step": ms_total / k, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": {L.U8: "u8", L.U16: "u16", L.F32: "f32"}[w["dtype"]] + "->f32",
            "data": "synthetic", "gsamples_per_s": gsps,
            "config": {"workload": w["label"], "volume": ("procedural " if w.get("procedural") else "") + ("V_noise seed 0x5EED" if w["kind"] == 1 else "V_sph"),
                       "camera": "36-step orbit (Ry 10deg steps, Rx 20deg), eye (0,0,1.6) fov 50",
                       "parallelism": par,
                       "path": ("HQ MIP frame (per-brick GLRaycaster-MIP-Rot-FS + Transfer-MIP, PlanHQMIPFrame LoD)" if mip else
                                "classic per-brick GLRaycaster") if classic else "GridLeaper page-table traversal",
                       "l2_policy": "inputs larger than L2 (pool %.1f GB logical, %.0f MB of bricks touched per frame)" %
                                    (info.pool_capacity[0] * info.pool_capacity[1] * info.pool_capacity[2] * slot_bytes / 1e9,
                                     step_touched / k * slot_bytes / 1e6),
                       "max_gradient_magnitude": 0.25, "hash_table": "collision-free (one slot per brick id; reference default 509)",
                       "timed_frames_not_converged": int(not_conv_max),
                       "bricks_paged_in_setup": paged, "setup_s": round(setup_s, 2),
                       "samples_per_frame": step_samples / k, "rays_per_frame": float(np.mean(rays)),
                       "bricks_touched_per_frame": step_touched / k,
                       "lane_utilisation": None if classic else
                                           {"sampling": float(np.sum(samples)) / max(1.0, 32.0 * warp_it),
                                            "alive": alive_it / max(1.0, 32.0 * warp_it)}},
            "parity": parity,
            "roofline": {"bound": "latency/issue" if not classic else "hbm", "contract_bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which, "kernel": "classic_kernel" if classic else "raycast_kernel",
                         "kernel_ms": ray_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "limiter": "latency / issue: dependent arithmetic at 4 resident warps per scheduler (DESIGN.md section 3); "
                                    "the HBM fraction is reported as the contract asks, the fetch-path fraction is the one that "
                                    "measures this kernel",
                         "fetch": fetch},
            "e2e": {"value": k / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(C_sizeof_params()),
                    "d2h_bytes_per_step": n_pixels * 4 + 8},
            # kernels of this library launched inside the timed region, per step: traversal + miss-table compaction (+ the LPT
            # tile sort, on unless TVK_TILE_LPT=0); library sort-last per rank in addition: wait-consumed, signal, wait-ready
            # and the n-way blend (+ rank 0's wait for the gathered slices); memsets / copies are not kernels and not counted
            "gpu_launches": k * (2 * world if pipe is not None else
                                 ((5 if split == "paired" else 2 + lpt_on + (4 if sl_mode[0] else 1)) * world + (1 if sl_mode[0] else 0)) if lib_sl else
                                 (2 + (0 if classic else lpt_on)) if sl is None else 2 + int(np.log2(world))),
            "clocks": clocks,
        }
        if stream_obj is not None:
            line["out_of_core"] = dict(PROC_INFO, cold_start=cold, streaming=stream_obj)
        if composite is not None:
            line["parity_composite"] = composite
        if per_rank is not None:
            line["per_rank"] = {"columns": ["kernel_ms", "exchange_ms (slices + blend + gather, incl. waiting for the slowest rank)",
                                            "frame_ms", "of exchange_ms: waiting for the peers' partial images (peer-memory path)",
                                            "samples_per_frame", "peak_device_GiB"], "rows": per_rank}
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            cs = CpuSample(args.config, args.cpu_vol, threads)
            best, n_s = min(cs.frame() for _ in range(3))
            cpu_sps = n_s / best
            line["cpu_baseline"] = {"value": cpu_sps / (step_samples / k), "unit": "frames/s", "cores": threads,
                                    "kind": cs.kind, "gsamples_per_s": cpu_sps / 1e9,
                                    "sample": cs.desc + "; value = CPU samples/s / samples of one GPU frame"}
            if cs.kind == "reference":
                pt, ps = min(cs.frame_port() for _ in range(2))
                line["cpu_baseline"]["port_gsamples_per_s"] = ps / pt / 1e9
                line["cpu_baseline"]["port_value"] = (ps / pt) / (step_samples / k)
        print(json.dumps(line), flush=True)
        if (parity is not None and not parity["ok"]) or (composite is not None and not composite["ok"]):
            rc = 3      # the reported frames are outside the parity gate: the number does not count
    r.Cleanup()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def C_sizeof_params():
    import ctypes
    from tuvok_b200 import _lib as L
    return ctypes.sizeof(L.RenderParams)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    run_tvk(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
