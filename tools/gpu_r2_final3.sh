#!/bin/bash
# record run of the round's final code on one GPU: smoke, GPU tests, the driver's bench command, the reference arm
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
echo "launch list rc=$?"
