"""Multi-GPU robustness check of the sort-last frame with the overlapped exchange (run under torchrun, one rank per GPU):
a small volume, every frame of a walk through different views READ BACK and compared with the single-GPU frame of the same
view (a frame delivered one exchange late, or torn by the next frame's gather, fails), then the same after a window resize
(buffers re-allocated, peers re-mapped), after a transfer-function change, with the asynchronous read-back, and after a
shutdown / re-init of the communicator.  Prints one line per phase; exit code 1 on any mismatch.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/gpu_sl_robust.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import tuvok_b200 as tb  # noqa: E402
from tuvok_b200 import _lib as L, sortlast, synth, workloads  # noqa: E402
import parity_gate  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
SIZE, BRICK, OV = (256, 256, 256), 36, 2
vol = synth.synth_volume(synth.V_NOISE, SIZE, L.U16, 0x5EED)
w = dict(workloads.WORKLOADS["c3"]); w["size"] = SIZE
t1, t2 = workloads.transfer_functions(w)


def make(width, height, shard=None):
    r = tb.CudaGridLeaper(device=local, max_gpu_mem=4 << 30, hash_table_size=8 * 8 * 8 * 4 + 64)
    if shard is not None:
        r.SetStoreShard(*shard)
    r.BuildVolume(vol, BRICK, OV, max_gradient_magnitude=0.25)
    r.Set1DTrans(t1); r.Set2DTrans(t2); r.SetRendermode(L.RM_2DTRANS); r.SetUseLighting(True)
    r.Resize(width, height); r.CreateVolumePool()
    return r


def single_frames(views, width, height, tf2=None):
    out = {}
    r0 = make(width, height)
    if tf2 is not None:
        r0.Set2DTrans(tf2)
    for v in views:
        r0.SetRotation(workloads.orbit_rotation(v, 36))
        assert r0.PaintUntilConverged().converged
        r0._dirty = True; r0.Paint()
        out[v] = r0.ReadRGBA8().copy()
    r0.Cleanup()
    return out


def all_min(flag):
    t = torch.tensor([1.0 if flag else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item()) > 0.5


def converge(r, v):
    for _ in range(64):
        r.SetRotation(workloads.orbit_rotation(v, 36))
        if all_min(r.SortLastFrame().frame.converged):
            return
    raise RuntimeError("view %d did not converge" % v)


bad = 0


def check(tag, got, ref):
    global bad
    mx, psnr = parity_gate.image_metrics(got, ref)
    over2 = int((np.abs(got.astype(np.int32) - ref.astype(np.int32)).max(axis=-1) > 2).sum())
    ok = psnr >= 45.0 and over2 <= max(2, int(1e-5 * got.shape[0] * got.shape[1]))
    bad += 0 if ok else 1
    print("%-44s max %d/255  psnr %.1f dB  pixels > 2: %d  %s" % (tag, mx, psnr, over2, "ok" if ok else "MISMATCH"), flush=True)


inner = BRICK - 2 * OV
finest = [-(-v // inner) for v in SIZE]
fl = [np.float32(v) / np.float32(inner) for v in SIZE]
fl = [f - f * np.finfo(np.float32).eps for f in fl]
ext = (1.0, 1.0, 1.0)
cmin, cmax, _ = sortlast.plan(finest, fl, ext, np.eye(4, dtype=np.float32) + 0, world, L.SL_OCTANT)
shard = (tuple(float(v) for v in cmin[rank]), tuple(float(v) for v in cmax[rank]))
VIEWS = [0, 7, 13, 22, 31, 4]
for phase, (W, H) in enumerate([(640, 360), (512, 512)]):
    ref = single_frames(VIEWS, W, H) if rank == 0 else None
    dist.barrier()
    if phase == 0:
        r = make(W, H, shard)
        sortlast.init_library_sortlast(r, rank, world, dist, L.SL_OCTANT)
    else:
        r.Resize(W, H)                                   # frame buffers, published image and peer mappings are renewed
    for v in VIEWS:                                      # page everything in once
        converge(r, v)
    # every frame of a walk read back synchronously, one frame per view
    for v in VIEWS + VIEWS[::-1]:
        r.SetRotation(workloads.orbit_rotation(v, 36))
        st = r.SortLastFrame()
        if rank == 0:
            check("%dx%d view %2d (sync read, mode %d)" % (W, H, v, st.peer_memory), r.SortLastReadRGBA8(), ref[v])
    # asynchronous read-back, two frames in flight: frame i is checked after frame i + 1 has been queued
    if rank == 0:
        pinned = [r.host_alloc((H, W, 4)) for _ in range(2)]
    prev = None
    for i, v in enumerate(VIEWS):
        r.SetRotation(workloads.orbit_rotation(v, 36))
        r.SortLastFrame()
        if rank == 0:
            r.SortLastReadRGBA8Async(pinned[i % 2])
            r.WaitRead(pending_allowed=1)
            if prev is not None:
                check("%dx%d view %2d (async read)" % (W, H, prev[1]), pinned[prev[0]].copy(), ref[prev[1]])
            prev = (i % 2, v)
    if rank == 0:
        r.WaitRead(0)
        check("%dx%d view %2d (async read, last)" % (W, H, prev[1]), pinned[prev[0]].copy(), ref[prev[1]])
    dist.barrier()
# a transfer-function change between two frames (visibility recomputed, pool re-paged)
from tuvok_b200.tf import TransferFunction2D  # noqa: E402
tf_b = TransferFunction2D.rectangle(w=t2.GetSize()[0], h=t2.GetSize()[1], x0=0.3, x1=0.8, alpha_max=40)
ref = single_frames([5, 17], 512, 512, tf_b) if rank == 0 else None
dist.barrier()
r.Set2DTrans(tf_b)
for v in (5, 17):
    converge(r, v)
    r.SetRotation(workloads.orbit_rotation(v, 36))
    r.SortLastFrame()
    if rank == 0:
        check("512x512 view %2d (after a TF change)" % v, r.SortLastReadRGBA8(), ref[v])
# shutdown and re-init of the sort-last state on the same renderer
r.SortLastShutdown()
dist.barrier()
sortlast.init_library_sortlast(r, rank, world, dist, L.SL_OCTANT)
for v in (5, 17):
    converge(r, v)
    r.SetRotation(workloads.orbit_rotation(v, 36))
    r.SortLastFrame()
    if rank == 0:
        check("512x512 view %2d (after shutdown / re-init)" % v, r.SortLastReadRGBA8(), ref[v])
dist.barrier()
r.SortLastShutdown()
r.Cleanup()
flag = torch.tensor([float(bad)], device="cuda")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
if rank == 0:
    print("RESULT:", "all frames match" if bad == 0 else "%d MISMATCHES" % bad, flush=True)
sys.exit(1 if flag.item() > 0 else 0)
