"""Developer probe (gpurun --gpus N, one process): peer-to-peer copy bandwidth and latency between GPU 0 and the others."""
import torch, subprocess
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
n = torch.cuda.device_count()
for p in range(1, n):
    print("can_access 0->%d" % p, torch.cuda.can_device_access_peer(0, p))
    for mb in (1, 16, 256):
        a = torch.empty(mb << 20, dtype=torch.uint8, device="cuda:0")
        b = torch.empty(mb << 20, dtype=torch.uint8, device="cuda:%d" % p)
        torch.cuda.set_device(0)
        b.copy_(a); torch.cuda.synchronize(0); torch.cuda.synchronize(p)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            a.copy_(b)          # device 0 pulls from the peer
        e.record(); torch.cuda.synchronize(0)
        print("  pull %4d MiB from GPU %d: %.1f GB/s (%.1f us per copy)" % (mb, p, 10 * (mb << 20) / 1e9 / (s.elapsed_time(e) * 1e-3), s.elapsed_time(e) * 100))
