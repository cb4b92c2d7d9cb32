#!/bin/bash
for m in DCS ECS; do
timeout -s KILL 120 python tools/prof_run.py $m 3e5 3 > gpurun_out/plain_$m.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_${m,,}_ -s 2 -c 2 -o gpurun_out/prof_${m,,}_r1 python tools/prof_run.py $m 3e5 3 > gpurun_out/ncu_$m.log 2>&1
tail -2 gpurun_out/plain_$m.log
done
