#!/bin/bash
timeout -s KILL 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for l in 1e6 1e7; do for cap in 64 256 1024; do
 echo "== l=$l cap=$cap"; CAP=$cap timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1
done; done
