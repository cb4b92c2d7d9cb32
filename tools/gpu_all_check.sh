#!/bin/bash
timeout -s KILL 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for m in MHRS:1e7 DCS:4e6 ECS:4e6; do M=${m%%:*}; L=${m##*:}; timeout -s KILL 300 python tools/prof_run.py $M $L 3 2>&1 | tail -1 | cut -c1-330; done
