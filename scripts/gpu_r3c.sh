#!/bin/bash
# ncu on the launches of two emulated ranks of the 8-way octant partition (light: rank 0, heavy: rank 3), view 0
P=${1:-r3c}
mkdir -p gpurun_out
for g in 0 3; do
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -c 12 -o gpurun_out/${P}_rank$g \
    python scripts/gpu_shard_probe.py --n 8 --split octant --views 0 --only-rank $g --repeat 2 > gpurun_out/${P}_rank$g.log 2>&1
tail -2 gpurun_out/${P}_rank$g.log
done
python scripts/gpu_shard_probe.py --n 8 --split octant --views 0,9 --repeat 4 2>&1 | tail -20 | tee gpurun_out/${P}_probe.txt
ls -la gpurun_out/${P}_rank*.ncu-rep
