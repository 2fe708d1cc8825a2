#!/bin/bash
# 8-GPU box: library sort-last (peer memory vs NCCL) on C3 and the out-of-core config C5
P=${1:-r2q}; N=${2:-8}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
nproc; free -g | head -2 | tail -1
run() {  # name, env, args
  name=$1; shift
  timeout 900 env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N ${@:2} > gpurun_out/${P}_$name.json 2> gpurun_out/${P}_$name.err
  echo "$name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_$name.json").read().strip().splitlines()[-1])
    print("$name fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","float_bit_identical","pixels")})
    c=d.get("parity_composite"); print("composite", {k: c[k] for k in c if k not in ("checker","worst")}); print("per_rank", d.get("per_rank")["rows"])
    print(d["config"]["parallelism"][:140])
    if "out_of_core" in d: print("ooc", json.dumps(d["out_of_core"]))
except Exception as e:
    print("bench parse failed", e)
PY
  grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/${P}_$name.err | tail -4
}
run c3_n${N}_peer1 TVK_SL_PEER=1 --steps 108 --warmup 4 --split octant
run c3_n${N}_peer0 TVK_SL_PEER=0 --steps 108 --warmup 4 --split octant
run c5_n${N} TVK_SL_PEER=1 --config c5 --steps 72 --warmup 4 --split octant
