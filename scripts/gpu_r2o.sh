#!/bin/bash
P=${1:-r2o}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_data.py -m gpu -q 2>&1 | tail -8
TVK_BUILD_TRACE=1 python scripts/bricker_time.py > gpurun_out/${P}_bricker.txt 2>&1; cat gpurun_out/${P}_bricker.txt
