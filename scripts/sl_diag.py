"""Developer tool (torchrun under gpurun --gpus N): per-view, per-rank samples and kernel ms of the sort-last path."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import tuvok_b200 as tb
from tuvok_b200 import _lib as L, sortlast, workloads
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", lr)); torch.cuda.set_device(lr)
vd = int(os.environ.get("VIEW_DEP", "1"))
w = dict(workloads.WORKLOADS["c3"])
inner = w["brick"] - 2 * w["overlap"]; finest = [-(-v // inner) for v in w["size"]]
r = tb.CudaGridLeaper(device=lr, max_gpu_mem=96 << 30, hash_table_size=finest[0] * finest[1] * finest[2] * 7 + 8)
raw = torch.empty(2 * 2048 ** 3, dtype=torch.uint8, device="cuda")
r.synth_volume(raw.data_ptr(), w["kind"], w["size"], w["dtype"], 0x5EED)
r.BuildVolume(raw.data_ptr(), w["brick"], w["overlap"], size=w["size"], dtype=w["dtype"], max_gradient_magnitude=0.25)
del raw; torch.cuda.empty_cache()
t1, t2 = workloads.transfer_functions(w); r.Set1DTrans(t1); r.Set2DTrans(t2); r.SetRendermode(w["mode"]); r.SetUseLighting(True)
r.Resize(w["width"], w["height"]); r.CreateVolumePool()
fl = [np.float32(v) / np.float32(inner) for v in w["size"]]; fl = [f - f * np.finfo(np.float32).eps for f in fl]
sl = sortlast.SortLastRenderer(r, rank, world, finest, fl, (1.0, 1.0, 1.0), view_dependent=bool(vd), policy=os.environ.get("SPLIT", "screen"))
if os.environ.get("SPLIT", "screen").endswith("w"):
    r.SetRotation(workloads.orbit_rotation(0)); r.Paint(); sl.set_weights()
VIEWS = int(os.environ.get("VIEWS", "36"))
rows = []
for i in range(36):
    r.SetRotation(workloads.orbit_rotation(i)); sl.update_partition(); r.PaintUntilConverged()
r.enable_counters(True)
for i in range(36):
    r.SetRotation(workloads.orbit_rotation(i)); sl.update_partition(); st = r.Paint(); rows.append([st.samples, st.rays, 0.0])
r.enable_counters(False)
for i in range(36):
    r.SetRotation(workloads.orbit_rotation(i)); sl.update_partition(); st = r.Paint(); rows[i][2] = st.ms_raycast; rows[i].append(str(sl._axes))
allr = [None] * world
dist.all_gather_object(allr, rows)
if rank == 0:
    for i in range(36):
        print("view %2d axes %s" % (i, allr[0][i][3]), " | ".join("%5.1fM smp %6.0fk rays %5.2f ms" % (a[i][0] / 1e6, a[i][1] / 1e3, a[i][2]) for a in allr))
    print("sum of per-view max ms: %.2f; per-rank sums: %s" % (sum(max(a[i][2] for a in allr) for i in range(36)), [round(sum(a[i][2] for i in range(36)), 2) for a in allr]))
dist.destroy_process_group()
