#!/bin/bash
timeout -s KILL 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for lib in libpht_mb2.so libpht_b200.so libpht_mb4.so; do
  for l in 1e6 1e7; do
    echo "== $lib l=$l"; PHT_B200_LIB=$PWD/phasetype_b200/$lib timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1
  done
done
