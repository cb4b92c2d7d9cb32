#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_mhrs_gpu.py tests/test_edges_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py tests/test_multigpu_gpu.py -q -m gpu -x 2>&1 | tail -3
for l in 1.25e6 1e7; do echo -n "l=$l: "; timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'attempts': ([0-9]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 attempts \2 ns_lane \3 ns_tail \4 ns_replay \5/"; done
TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 6 2>&1 | tail -12
