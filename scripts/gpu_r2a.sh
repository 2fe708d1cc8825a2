#!/bin/bash
# round 2, first GPU pass: tree state after the advisor fixes + the full-size parity gate on the round-1 kernel
P=r2a
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${P}_pytest.log
tail -3 gpurun_out/${P}_pytest.log
for c in c3 c2 c4; do
  timeout 600 python scripts/parity_full.py --config $c --stride 8 > gpurun_out/${P}_parity_$c.json 2> gpurun_out/${P}_parity_$c.err
  echo "parity $c rc=$?"; cat gpurun_out/${P}_parity_$c.json; tail -3 gpurun_out/${P}_parity_$c.err
done
timeout 600 python bench.py --steps 216 --warmup 4 --no-cpu > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
cat gpurun_out/${P}_bench_c3.json; tail -3 gpurun_out/${P}_bench_c3.err
