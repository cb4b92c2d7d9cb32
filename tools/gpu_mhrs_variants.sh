#!/bin/bash
echo default; timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-90
for v in m224 m256x4; do echo $v; PHT_B200_LIB=$PWD/phasetype_b200/libpht_$v.so timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-90; done
