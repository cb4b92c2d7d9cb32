#!/bin/bash
# round-2 record run on one GPU: all GPU tests, the default bench line, the e2e breakdown
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
timeout -s KILL 300 python tools/e2e_breakdown.py 2>&1 | tail -4
