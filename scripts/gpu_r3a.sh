#!/bin/bash
# LPT tile schedule A/B on one GPU: parity tests with the schedule on, then C3 / C3t with TVK_TILE_LPT = 0 / 1
P=${1:-r3a}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
TVK_TILE_LPT=1 timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_sortlast.py tests/test_parity_gate.py -m gpu -q -x 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -6
for cfg in c3 c3t; do
for lpt in 0 1; do
TVK_TILE_LPT=$lpt timeout 600 python bench.py --config $cfg --steps 108 --no-cpu 2> gpurun_out/${P}_${cfg}_lpt$lpt.err | tail -1 > gpurun_out/${P}_${cfg}_lpt$lpt.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_${cfg}_lpt$lpt.json").read().strip().splitlines()[-1])
    print("$cfg lpt=$lpt fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f parity %s" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"], d["parity"]["float_bit_identical"]))
except Exception as e:
    print("bench parse failed", e)
PY
done
done 2>&1 | tee gpurun_out/${P}_lpt_ab.txt
