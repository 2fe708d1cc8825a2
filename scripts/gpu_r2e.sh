#!/bin/bash
P=${1:-r2e}
mkdir -p gpurun_out
bash scripts/ab_variants.sh 2>&1 | tee gpurun_out/${P}_ab.log
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 40 -c 2 -o gpurun_out/${P}_prof \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${P}_ncu_full.log 2>&1
ls -la gpurun_out/${P}_prof.ncu-rep
