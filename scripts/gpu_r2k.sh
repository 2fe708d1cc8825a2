#!/bin/bash
# 1 GPU: c5 (8192^3 u8 procedural, out-of-core) bench line; optional camera variants TZ list
P=${1:-r2k}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
for tz in ${2:-default}; do
if [ "$tz" != "default" ]; then export TVK_TZ=$tz; fi
timeout 1500 python bench.py --config c5 --steps 72 --warmup 4 > gpurun_out/${P}_bench_c5_$tz.json 2> gpurun_out/${P}_bench_c5_$tz.err
echo "tz=$tz rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_bench_c5_$tz.json").read().strip().splitlines()[-1])
    print("c5 fps %.1f e2e %.1f gsps %.2f kernel_ms %.3f" % (d["value"], d["e2e"]["value"], d["gsamples_per_s"], d["roofline"]["kernel_ms"]))
    print("parity", {k: d["parity"][k] for k in ("ok","max_abs_255","float_bit_identical","pixels")}); print("ooc", json.dumps(d.get("out_of_core"))); print("cfg", {k: d["config"][k] for k in ("bricks_paged_in_setup","setup_s","samples_per_frame","bricks_touched_per_frame","l2_policy")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/${P}_bench_c5_$tz.err
done
