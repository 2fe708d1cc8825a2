#!/bin/bash
# Short GPU-box pass (run under gpurun): GPU parity tests, smoke, the C3 bench line (with the CPU baseline), the
# reference arm and the ncu launch list.  Outputs land in gpurun_out/ with prefix $1.
P=${1:-r1}
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${P}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
python bench.py > gpurun_out/${P}_bench_c3.json 2> gpurun_out/${P}_bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${P}_ncu_bench.log 2>&1
tail -3 gpurun_out/${P}_pytest.log; tail -2 gpurun_out/${P}_smoke.log; cat gpurun_out/${P}_bench_ref.json gpurun_out/${P}_bench_c3.json
