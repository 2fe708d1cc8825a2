#!/bin/bash
P=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${P}_pytest.log
tail -4 gpurun_out/${P}_pytest.log
timeout 600 python scripts/parity_full.py --config c3 --stride 8 --views 0,12 > gpurun_out/${P}_parity_c3.json 2> gpurun_out/${P}_parity_c3.err
echo "parity c3 rc=$?"; cut -c1-120 gpurun_out/${P}_parity_c3.json
bash scripts/ab_variants.sh 2>&1 | tee gpurun_out/${P}_ab.log
bash scripts/ab_variants.sh --config c2 2>&1 | tee gpurun_out/${P}_ab_c2.log
bash scripts/ab_variants.sh --config c4 2>&1 | tee gpurun_out/${P}_ab_c4.log
