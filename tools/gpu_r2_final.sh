#!/bin/bash
# round-2 record run on one GPU: GPU tests, the driver's bench command, ncu launch list of the bench, full captures of the path kernels
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
echo "launch list rc=$?"
cap() { # METHOD L TAG REGEX SKIP COUNT [GENERAL]
  GENERAL=$7 timeout -s KILL 300 python tools/prof_run.py $1 $2 3 > gpurun_out/plain_$3.log 2>&1 && \
  GENERAL=$7 timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:$4 -s $5 -c $6 -f -o gpurun_out/prof_$3 python tools/prof_run.py $1 $2 3 > gpurun_out/ncu_$3.log 2>&1
  echo "capture $3 rc=$?"; tail -1 gpurun_out/plain_$3.log | cut -c1-120
}
cap MHRS 1e7 mhrs_r2f k_mhrs_ 3 3 ""
cap DCS 4e6 dcs_r2f k_dcs_sweep 2 2 1
cap ECS 4e6 ecs_r2f k_ecs_ 2 2 1
ls -la gpurun_out/*.ncu-rep | tail -4
