#!/bin/bash
timeout -s KILL 800 python -m pytest tests/test_mhrs_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py tests/test_edges_gpu.py tests/test_tiers_gpu.py -x -q -m gpu --durations=3 2>&1 | tail -7
for l in 1e6 1e7; do echo "l=$l"; timeout -s KILL 200 python tools/prof_run.py MHRS $l 5 2>&1 | tail -1 | cut -c1-90; done
