#!/bin/bash
# fresh MHRS profile of the current kernel (4e6 observations, sweep 2) + plain timings at 1e6/1e7
timeout -s KILL 200 python tools/prof_run.py MHRS 4e6 3 > gpurun_out/plain_mhrs4e6.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_mhrs_sweep -s 1 -c 1 -o gpurun_out/prof_mhrs_r1c python tools/prof_run.py MHRS 4e6 3 > gpurun_out/ncu_mhrs4e6.log 2>&1
tail -2 gpurun_out/plain_mhrs4e6.log; tail -3 gpurun_out/ncu_mhrs4e6.log
timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 > gpurun_out/plain_mhrs1e7.log 2>&1; tail -1 gpurun_out/plain_mhrs1e7.log
