#!/bin/bash
# 1 GPU, final evidence of round 2: GPU suite, smoke, bench lines (c3 + reference arm, c3t, c2, c4), ncu launch list and one
# full capture of the traversal kernel (summarised on the box: the report itself exceeds the 64 MiB return limit)
P=${1:-r3f}
mkdir -p gpurun_out
make -C oracle liborc.so > /dev/null 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 > gpurun_out/${P}_pytest_full.log
grep -v "warning\|orc_render.c\|^\s*[0-9]* |\|string_fortified\|~~\|In function\|inlined\|In file\|from \|^\s*|" gpurun_out/${P}_pytest_full.log | tail -30 > gpurun_out/${P}_pytest.log
rm -f gpurun_out/${P}_pytest_full.log
tail -4 gpurun_out/${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err; echo "c3 rc=$?"
for c in c3t c2 c4; do timeout 600 python bench.py --config $c --no-cpu > gpurun_out/${P}_bench_$c.json 2> gpurun_out/${P}_bench_$c.err; echo "$c rc=$?"; done
python - <<PY
import json
for c in ("c3","ref","c3t","c2","c4"):
    try:
        d=json.loads(open("gpurun_out/${P}_bench_%s.json"%c).read().strip().splitlines()[-1])
        print(c, "fps %.2f e2e %.2f gsps %.3f" % (d["value"], d["e2e"]["value"], d.get("gsamples_per_s",0)), "kernel_ms", d.get("roofline",{}).get("kernel_ms"), "fetch", (d.get("roofline",{}).get("fetch") or {}).get("frac"), "parity", (d.get("parity") or {}).get("max_abs_255"), (d.get("parity") or {}).get("float_bit_identical"))
    except Exception as e: print(c, "parse failed", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-parity > gpurun_out/${P}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 40 -c 1 -o /tmp/${P}_prof python bench.py --steps 4 --warmup 3 --no-cpu --no-parity > gpurun_out/${P}_ncu_full.log 2>&1
python scripts/ncu_summary.py /tmp/${P}_prof.ncu-rep > gpurun_out/${P}_raycast_c3_ncu_summary.txt 2>&1
head -24 gpurun_out/${P}_raycast_c3_ncu_summary.txt
python scripts/launch_summary.py gpurun_out/${P}_launches.csv > gpurun_out/${P}_launches_summary.txt 2>&1; tail -12 gpurun_out/${P}_launches_summary.txt
du -sh gpurun_out
