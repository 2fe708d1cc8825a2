#!/bin/bash
# One GPU-box pass (run under gpurun): GPU parity tests, smoke, the bench lines, the ncu launch list and
# one full ncu capture of the traversal kernel.  Outputs land in gpurun_out/ with prefix $1 (default r1).
P=${1:-r1}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1
python bench.py > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
python bench.py --config c2 --no-cpu > gpurun_out/${P}_bench_c2.json 2> gpurun_out/${P}_bench_c2.err
python bench.py --config c4 --no-cpu > gpurun_out/${P}_bench_c4.json 2> gpurun_out/${P}_bench_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${P}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 40 -c 2 -o gpurun_out/${P}_prof \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${P}_ncu_full.log 2>&1
tail -3 gpurun_out/${P}_pytest.log; tail -2 gpurun_out/${P}_smoke.log; cat gpurun_out/${P}_bench_c3.json
