#!/bin/bash
# N GPUs: library sort-last through peer memory, exchange overlapped with the next frame's traversal (1) vs in order (0)
P=${1:-r2w}; N=${2:-2}; MODES=${3:-"1 0"}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
for ov in $MODES; do
export TVK_SL_OVERLAP=$ov
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 108 --warmup 4 --split octant > gpurun_out/${P}_bench_n${N}_overlap$ov.json 2> gpurun_out/${P}_bench_n${N}_overlap$ov.err
echo "n=$N overlap=$ov rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_n${N}_overlap$ov.json").read().strip().splitlines()[-1])
    print("n=$N overlap=$ov fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","float_bit_identical","pixels")})
    c=d.get("parity_composite"); print("composite", {k: c[k] for k in c if k not in ("checker","worst")}); print("per_rank", d.get("per_rank")["rows"])
    print(d["config"]["parallelism"][-260:])
except Exception as e:
    print("bench parse failed", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/${P}_bench_n${N}_overlap$ov.err | tail -4
done
