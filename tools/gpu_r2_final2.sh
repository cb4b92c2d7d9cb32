#!/bin/bash
# last record run of round 2 on one GPU: smoke, GPU tests, the driver's bench command, launch list, ECS recapture (k_ecs_gt restructured)
mkdir -p gpurun_out
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/r2_bench_ref.json
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
echo "launch list rc=$?"
GENERAL=1 timeout -s KILL 300 python tools/prof_run.py ECS 4e6 3 > gpurun_out/plain_ecs_r2g.log 2>&1 && \
GENERAL=1 timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k_ecs_ -s 2 -c 2 -f -o gpurun_out/prof_ecs_r2g python tools/prof_run.py ECS 4e6 3 > gpurun_out/ncu_ecs_r2g.log 2>&1
echo "capture ecs rc=$?"; tail -1 gpurun_out/plain_ecs_r2g.log | cut -c1-100
