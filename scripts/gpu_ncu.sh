#!/bin/bash
# one full ncu capture of the traversal kernel on the C3 bench workload (+ the launch list)
P=${1:-r2c}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 40 -c 2 -o gpurun_out/${P}_prof \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${P}_ncu_full.log 2>&1
tail -3 gpurun_out/${P}_ncu_full.log
ls -la gpurun_out/${P}_prof.ncu-rep
