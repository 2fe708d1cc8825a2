#!/bin/bash
P=${1:-r2b}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${P}_pytest.log
tail -25 gpurun_out/${P}_pytest.log
for c in c3 c2 c4; do
  timeout 600 python scripts/parity_full.py --config $c --stride 8 --views 0,12 > gpurun_out/${P}_parity_$c.json 2> gpurun_out/${P}_parity_$c.err
  echo "parity $c rc=$?"; cut -c1-330 gpurun_out/${P}_parity_$c.json; tail -3 gpurun_out/${P}_parity_$c.err
done
timeout 600 python bench.py --steps 216 --warmup 4 --no-cpu > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
cut -c1-200 gpurun_out/${P}_bench_c3.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/${P}_bench_c3.json; grep -o '"e2e": {[^}]*}' gpurun_out/${P}_bench_c3.json; tail -3 gpurun_out/${P}_bench_c3.err
for c in c2 c4; do
timeout 600 python bench.py --config $c --steps 216 --warmup 4 --no-cpu > gpurun_out/${P}_bench_$c.json 2> gpurun_out/${P}_bench_$c.err
cut -c1-200 gpurun_out/${P}_bench_$c.json; grep -o '"kernel_ms": [0-9.]*' gpurun_out/${P}_bench_$c.json; tail -3 gpurun_out/${P}_bench_$c.err
done
