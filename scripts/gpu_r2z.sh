#!/bin/bash
P=${1:-r2z}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 900 python -m pytest tests/test_mip.py tests/test_classic.py -m gpu -q -x 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -40 | tee gpurun_out/${P}_mip_ortho.txt
