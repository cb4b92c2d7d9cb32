#!/bin/bash
# usage: gpu_prof4.sh METHOD L TAG [kernel regex]
M=$1; L=$2; TAG=$3; RX=${4:-k_${M,,}_}
timeout -s KILL 200 python tools/prof_run.py $M $L 3 > gpurun_out/plain_$TAG.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:$RX -s 1 -c 1 -o gpurun_out/prof_$TAG python tools/prof_run.py $M $L 3 > gpurun_out/ncu_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log; tail -2 gpurun_out/ncu_$TAG.log
