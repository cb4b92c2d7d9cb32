#!/bin/bash
timeout -s KILL 600 python -m pytest tests/test_dcs_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -2
echo default; timeout -s KILL 200 python tools/prof_run.py DCS 4e6 3 2>&1 | tail -1 | cut -c1-90
for v in w20u2 w24u2 w16u4 w20u4; do echo $v; PHT_B200_LIB=$PWD/phasetype_b200/libpht_$v.so timeout -s KILL 200 python tools/prof_run.py DCS 4e6 3 2>&1 | tail -1 | cut -c1-90; done
