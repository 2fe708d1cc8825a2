#!/bin/bash
# 1 GPU: the whole -m gpu suite (new: clip plane, pick)
P=${1:-r2i}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/${P}_pytest_full.log
grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" gpurun_out/${P}_pytest_full.log | tail -150 > gpurun_out/${P}_pytest.log
tail -15 gpurun_out/${P}_pytest.log
