#!/bin/bash
for m in DCS ECS; do
 timeout -s KILL 400 python bench.py --method $m --obs 1000000 --steps 3 --warmup 3 > gpurun_out/bench_$m.log 2>&1; echo "== $m rc=$?"; tail -c 2600 gpurun_out/bench_$m.log
done
