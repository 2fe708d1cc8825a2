#!/bin/bash
# 8-GPU box: the library sort-last bench at N = 4 and 8 (octant + screen), composite gate included
P=${1:-r2h}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
for split in octant screen; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 108 --warmup 4 --split $split > gpurun_out/${P}_bench_n${n}_$split.json 2> gpurun_out/${P}_bench_n${n}_$split.err
echo "n=$n $split rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_n${n}_$split.json").read().strip().splitlines()[-1])
    print("$split n=$n fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","psnr_db","float_bit_identical","pixels")})
    c=d.get("parity_composite"); print("composite", {k: c[k] for k in c if k != "checker"}); print("per_rank", d.get("per_rank")["rows"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/${P}_bench_n${n}_$split.err
done
done
