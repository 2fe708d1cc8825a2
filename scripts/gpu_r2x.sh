#!/bin/bash
# persistent traversal kernel with lane refill: GPU parity tests + C3 bench
P=${1:-r2x}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -15 | tee gpurun_out/${P}_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_c3.json").read().strip().splitlines()[-1])
    print("fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","float_bit_identical","pixels")})
    print(d["config"].get("lane_utilisation"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/${P}_bench_c3.err
