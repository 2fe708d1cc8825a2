#!/bin/bash
P=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${P}_pytest.log
tail -25 gpurun_out/${P}_pytest.log
timeout 900 python bench.py --steps 216 --warmup 4 > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_c3.json").read().strip().splitlines()[-1])
    print("c3 fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", d["parity"]); print("fetch", d["roofline"]["fetch"]); print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/${P}_bench_c3.err
