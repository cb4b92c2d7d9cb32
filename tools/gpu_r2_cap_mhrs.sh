#!/bin/bash
# full ncu capture of the three MHRS kernels of one sweep (10^7 observations) with the round's final code
timeout -s KILL 300 python tools/prof_run.py MHRS 1e7 3 > gpurun_out/plain_mhrs_r2h.log 2>&1 && \
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:k_mhrs_ -s 3 -c 3 -f -o gpurun_out/prof_mhrs_r2h python tools/prof_run.py MHRS 1e7 3 > gpurun_out/ncu_mhrs_r2h.log 2>&1
echo "capture rc=$?"; tail -1 gpurun_out/plain_mhrs_r2h.log | cut -c1-120
