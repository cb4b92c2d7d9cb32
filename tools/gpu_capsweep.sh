#!/bin/bash
for c in 32 64 128 256; do echo "CAP=$c"; for l in 1e6 1e7; do CAP=$c timeout -s KILL 200 python tools/prof_run.py MHRS $l 5 2>&1 | tail -1 | cut -c1-120; done; done
