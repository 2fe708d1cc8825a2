#!/bin/bash
# A/B the raycast kernel build variants on N GPUs (developer tool): scripts/ab_variants_n.sh N
N=${1:-2}
for so in tuvok_b200/libtvk_var_*.so tuvok_b200/libtvkcuda.so; do
  TVK_LIB=$PWD/$so python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 36 --warmup 4 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=$N %-40s fps %.1f  kernel_ms %.3f  gsamples/s %.2f  e2e %.1f' % ('$so', d['value'], d['roofline']['kernel_ms'], d['gsamples_per_s'], d['e2e']['value']))"
done
