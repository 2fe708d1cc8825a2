#!/bin/bash
# 8-GPU box: paired policy at N = 8 and 4
P=${1:-r2s}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
for N in 8 4; do
split=paired
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 108 --warmup 4 --split $split > gpurun_out/${P}_bench_n${N}_$split.json 2> gpurun_out/${P}_bench_n${N}_$split.err
echo "n=$N $split rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_n${N}_$split.json").read().strip().splitlines()[-1])
    print("n=$N $split fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f samples %.1fM" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"], d["config"]["samples_per_frame"]/1e6))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","float_bit_identical","pixels")})
    c=d.get("parity_composite"); print("composite", {k: c[k] for k in c if k not in ("checker","worst")}); print("per_rank", d.get("per_rank")["rows"])
except Exception as e:
    print("bench parse failed", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/${P}_bench_n${N}_$split.err | tail -4
done
