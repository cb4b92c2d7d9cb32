#!/bin/bash
for l in 1.25e6 1e7; do
echo "default l=$l"; timeout -s KILL 200 python tools/prof_run.py MHRS $l 5 2>&1 | tail -1 | cut -c1-90
for v in c1 c3; do echo "$v l=$l"; PHT_B200_LIB=$PWD/phasetype_b200/libpht_$v.so timeout -s KILL 200 python tools/prof_run.py MHRS $l 5 2>&1 | tail -1 | cut -c1-90; done
done
