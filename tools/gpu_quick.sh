#!/bin/bash
# usage: gpu_quick.sh METHOD L [SWEEPS]
for l in ${2//,/ }; do timeout -s KILL 200 python tools/prof_run.py $1 $l ${3:-5} 2>&1 | tail -1; done
