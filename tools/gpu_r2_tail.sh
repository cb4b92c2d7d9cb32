#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | tail -8
for v in libpht_b200.so libpht_lg2.so libpht_lg8.so; do
echo "== $v"
PHT_B200_LIB=$PWD/phasetype_b200/$v TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -17 | cut -c1-110 | head -1
PHT_B200_LIB=$PWD/phasetype_b200/$v TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1.25e6 5 2>&1 | tail -16 | cut -c1-420 | head -14
done
