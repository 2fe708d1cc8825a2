#!/bin/bash
# 1 GPU: data-path tests (TMA bricker, median), bricker timing with and without TMA, c5 bench with upload trace
P=${1:-r2n}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_data.py -m gpu -q 2>&1 | tail -8
python scripts/bricker_time.py > gpurun_out/${P}_bricker.txt 2>&1; cat gpurun_out/${P}_bricker.txt
