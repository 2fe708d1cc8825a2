#!/bin/bash
# last check of the round on one GPU: full GPU suite, smoke, default bench line, c5 (procedural out-of-core) short
P=${1:-r3l}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" | tail -6 | tee gpurun_out/${P}_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee -a gpurun_out/${P}_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --config c5 --steps 36 --warmup 4 --no-cpu --no-stream > gpurun_out/${P}_bench_c5.json 2> gpurun_out/${P}_bench_c5.err; echo "c5 rc=$?"
python - <<PY
import json
for c in ("c3","c5"):
    try:
        d=json.loads(open("gpurun_out/${P}_bench_%s.json"%c).read().strip().splitlines()[-1])
        print(c, "fps %.2f e2e %.2f gsps %.3f" % (d["value"], d["e2e"]["value"], d.get("gsamples_per_s",0)), "kernel_ms", d.get("roofline",{}).get("kernel_ms"), "parity", (d.get("parity") or {}).get("max_abs_255"), (d.get("parity") or {}).get("float_bit_identical"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e: print(c, "parse failed", e)
PY
tail -2 gpurun_out/${P}_bench_c5.err | cut -c1-200
