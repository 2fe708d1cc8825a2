#!/bin/bash
V=$PWD/tuvok_b200/libtvk_var_split.so
for cfg in "base:0:$PWD/tuvok_b200/libtvkcuda.so" "split0:0:$V" "split300:300:$V" "split600:600:$V"; do
n=${cfg%%:*}; r=${cfg#*:}; sc=${r%%:*}; so=${r#*:}
echo "== $n"
TVK_SPLIT_COST=$sc TVK_LIB=$so python scripts/gpu_shard_probe.py --n 1 --views 0 --repeat 4 2>&1 | tail -2
TVK_SPLIT_COST=$sc TVK_LIB=$so python scripts/gpu_shard_probe.py --n 2 --split octant --views 0 --repeat 4 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r3k_split_probe.txt
