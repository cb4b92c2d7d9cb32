#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED|PASSED" | tail -15
echo "== dbg 1e6"; PHT_B200_LIB=$PWD/phasetype_b200/libpht_dbg.so timeout -s KILL 200 python tools/prof_run.py MHRS 1e6 5 2>&1 | tail -1 | cut -c1-700
echo "== dbg 1e7"; PHT_B200_LIB=$PWD/phasetype_b200/libpht_dbg.so timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-700
echo "== dbg 1e7 nosort"; PHT_B200_NO_SORT=1 PHT_B200_LIB=$PWD/phasetype_b200/libpht_dbg.so timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-700
