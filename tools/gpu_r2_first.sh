#!/bin/bash
# round 2, first look: MHRS parity (incl. the deep-tail sweep), model test, then timed sweeps of build variants
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py tests/test_model_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -15
for v in libpht_b200.so libpht_v3.so libpht_v2.so libpht_v43.so; do
  echo "== $v"; PHT_B200_LIB=$PWD/phasetype_b200/$v timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-600
done
echo "== 1e6"; timeout -s KILL 200 python tools/prof_run.py MHRS 1e6 5 2>&1 | tail -1 | cut -c1-600
