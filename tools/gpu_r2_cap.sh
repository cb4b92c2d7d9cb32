#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | tail -15
for c in 256 1024 4096 16384; do
  echo "== cap $c"; CAP=$c timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-640
done
echo "== 1e6 cap 256";  CAP=256 timeout -s KILL 200 python tools/prof_run.py MHRS 1e6 5 2>&1 | tail -1 | cut -c1-640
echo "== 1e6 cap 4096"; CAP=4096 timeout -s KILL 200 python tools/prof_run.py MHRS 1e6 5 2>&1 | tail -1 | cut -c1-640
