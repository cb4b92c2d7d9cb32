#!/bin/bash
set -x
timeout -s KILL 120 python tools/prof_run.py MHRS 1e6 3 > gpurun_out/plain.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_mhrs_sweep -s 1 -c 1 -o gpurun_out/prof_mhrs_r1b python tools/prof_run.py MHRS 1e6 3 > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/plain.log; tail -5 gpurun_out/ncu.log
