#!/bin/bash
# first measurement pass: smoke, tests, short benches
set -x
timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -5
timeout -s KILL 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --obs 1000000 > gpurun_out/bench_1e6.log 2>&1; tail -c 3000 gpurun_out/bench_1e6.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1e7.log 2>&1; tail -c 3000 gpurun_out/bench_1e7.log
