#!/bin/bash
# split work units (persistent variant, TVK_SPLIT_COST = cost threshold in turns): parity tests, C3 bench, emulated 8-way ranks
P=${1:-r3j}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
V=$PWD/tuvok_b200/libtvk_var_split.so
TVK_SPLIT_COST=200 TVK_LIB=$V timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_sortlast.py tests/test_parity_gate.py -m gpu -q -x 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -8
timeout 300 python -m pytest tests/test_gpu_data.py -m gpu -q -k default_pool 2>&1 | tail -2
for cfg in "base:0:$PWD/tuvok_b200/libtvkcuda.so" "split0:0:$V" "split300:300:$V" "split150:150:$V"; do
n=${cfg%%:*}; r=${cfg#*:}; sc=${r%%:*}; so=${r#*:}
echo "== $n"
TVK_SPLIT_COST=$sc TVK_LIB=$so timeout 600 python bench.py --steps 108 --no-cpu 2> gpurun_out/${P}_$n.err | tail -1 > gpurun_out/${P}_$n.json
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_$n.json").read().strip().splitlines()[-1])
    print("c3 fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f parity %s" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"], d["parity"]["float_bit_identical"]))
except Exception as e:
    print("bench parse failed", e)
PY
tail -2 gpurun_out/${P}_$n.err | cut -c1-300
TVK_SPLIT_COST=$sc TVK_LIB=$so python scripts/gpu_shard_probe.py --n 8 --split octant --views 0 --repeat 4 2>&1 | tail -9
done 2>&1 | tee gpurun_out/${P}_split_ab.txt
