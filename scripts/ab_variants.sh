#!/bin/bash
# A/B the raycast kernel build variants (developer tool): prints fps / kernel ms per variant
for so in tuvok_b200/libtvk_var_*.so tuvok_b200/libtvkcuda.so; do
  TVK_LIB=$PWD/$so python bench.py --steps 36 --warmup 4 --no-cpu "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-40s fps %.1f  kernel_ms %.3f  gsamples/s %.2f  e2e %.1f' % ('$so', d['value'], d['roofline']['kernel_ms'], d['gsamples_per_s'], d['e2e']['value']))"
done
