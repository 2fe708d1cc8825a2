#!/bin/bash
P=${1:-r2l}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_procedural.py -m gpu -q -x 2>&1 | tail -3
bash scripts/gpu_r2k.sh $P default
