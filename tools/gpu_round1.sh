#!/bin/bash
# round-1 evidence run: smoke, GPU tests, default bench, launch list of the bench command, full capture of each path kernel
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -s KILL 900 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r1.err
BCMD="python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e"
timeout -s KILL 600 $BCMD > gpurun_out/bench_short.json 2>&1 && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for m in MHRS DCS ECS; do
  timeout -s KILL 300 python tools/prof_run.py $m 1e7 3 > gpurun_out/plain_${m}_1e7.log 2>&1 && \
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_${m,,}_ -s 2 -c 2 -o gpurun_out/prof_${m,,}_r1_1e7 python tools/prof_run.py $m 1e7 3 > gpurun_out/ncu_${m}_1e7.log 2>&1
  echo "$m capture rc=$?"; tail -1 gpurun_out/plain_${m}_1e7.log | cut -c1-120
done
