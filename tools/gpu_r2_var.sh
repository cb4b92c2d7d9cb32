#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py -q -m gpu 2>&1 | grep -E "passed|failed|FAILED" | tail -8
for v in libpht_b200.so libpht_v3.so libpht_dbg.so; do
  echo "== $v"; PHT_B200_LIB=$PWD/phasetype_b200/$v timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 5 2>&1 | tail -1 | cut -c1-700
done
echo "== 1e6"; timeout -s KILL 200 python tools/prof_run.py MHRS 1e6 5 2>&1 | tail -1 | cut -c1-100
timeout -s KILL 300 python tools/prof_run.py MHRS 1e7 3 > gpurun_out/plain_mhrs_r2c.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_mhrs_ -s 2 -c 2 -o gpurun_out/prof_mhrs_r2c python tools/prof_run.py MHRS 1e7 3 > gpurun_out/ncu_mhrs_r2c.log 2>&1
echo "capture rc=$?"
