#!/bin/bash
# helper lanes (persistent variant + TVK_HELP): parity tests, C3 bench, emulated 8-way ranks; base library for comparison
P=${1:-r3e}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
TVK_LIB=$PWD/tuvok_b200/libtvk_var_h3.so timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_sortlast.py tests/test_parity_gate.py -m gpu -q -x 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -8
for so in libtvkcuda.so libtvk_var_h3.so libtvk_var_h7.so; do
echo "== $so"
TVK_LIB=$PWD/tuvok_b200/$so timeout 600 python bench.py --steps 108 --no-cpu 2> gpurun_out/${P}_$so.err | tail -1 > gpurun_out/${P}_$so.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_$so.json").read().strip().splitlines()[-1])
    print("c3 fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f parity %s" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"], d["parity"]["float_bit_identical"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -2 gpurun_out/${P}_$so.err | cut -c1-300
TVK_LIB=$PWD/tuvok_b200/$so python scripts/gpu_shard_probe.py --n 8 --split octant --views 0 --repeat 4 2>&1 | tail -9
done 2>&1 | tee gpurun_out/${P}_help_ab.txt
