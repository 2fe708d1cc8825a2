#!/bin/bash
# 1 GPU: procedural source tests + host probe (cores, memory, pinned H2D bandwidth)
P=${1:-r2j}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
nproc; free -g | head -2
python - <<PY
import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for n in (2 << 20, 32 << 20, 1 << 30):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d[:n].copy_(x[:n], non_blocking=True); torch.cuda.synchronize()
    s.record()
    for _ in range(8): d[:n].copy_(x[:n], non_blocking=True)
    e.record(); torch.cuda.synchronize()
    print("pinned H2D %4d MiB: %.1f GB/s" % (n >> 20, 8 * n / (s.elapsed_time(e) * 1e-3) / 1e9))
PY
timeout 900 python -m pytest tests/test_gpu_procedural.py -m gpu -q -x 2>&1 > gpurun_out/${P}_pytest_full.log
grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" gpurun_out/${P}_pytest_full.log | tail -60
