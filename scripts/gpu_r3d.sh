#!/bin/bash
# emulated ranks of the 8-way octant partition on one GPU: base library vs the software-prefetch variant
P=${1:-r3d}
mkdir -p gpurun_out
for so in tuvok_b200/libtvkcuda.so tuvok_b200/libtvk_var_prefetch.so; do
echo "== $so"
TVK_LIB=$PWD/$so python scripts/gpu_shard_probe.py --n 8 --split octant --views 0 --repeat 4 2>&1 | tail -9
done 2>&1 | tee gpurun_out/${P}_prefetch_probe.txt
