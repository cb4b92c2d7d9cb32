#!/bin/bash
timeout -s KILL 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo default; timeout -s KILL 200 python tools/prof_run.py ECS 4e6 3 2>&1 | tail -1 | cut -c1-90
for v in e20 e24; do echo $v; PHT_B200_LIB=$PWD/phasetype_b200/libpht_$v.so timeout -s KILL 200 python tools/prof_run.py ECS 4e6 3 2>&1 | tail -1 | cut -c1-90; done
