#!/bin/bash
# 2-GPU pass: GPU tests on one device, then the sort-last bench at N = 2 (library path, octant + screen)
P=${1:-r2g}
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_sortlast.py -m gpu -x -q 2>&1 | tail -8
for split in octant; do
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 108 --warmup 4 --split $split > gpurun_out/${P}_bench_n2_$split.json 2> gpurun_out/${P}_bench_n2_$split.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_n2_$split.json").read().strip().splitlines()[-1])
    print("$split n=2 fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","psnr_db","float_bit_identical","pixels")})
    print("composite", d.get("parity_composite")); print("per_rank", d.get("per_rank"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -6 gpurun_out/${P}_bench_n2_$split.err
done
