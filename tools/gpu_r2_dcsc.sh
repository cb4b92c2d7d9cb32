#!/bin/bash
# complex-spectrum DCS unit machine: tier-2 tests, then the sweep time on the general C3
timeout -s KILL 900 python -m pytest tests/test_complex_gpu.py tests/test_dcs_gpu.py tests/test_variants_gpu.py tests/test_tiers_gpu.py -q -m gpu -x 2>&1 | tail -4
GENERAL=1 timeout -s KILL 300 python tools/prof_run.py DCS 1e7 3 2>&1 | tail -1 | cut -c1-400
timeout -s KILL 300 python tools/prof_run.py DCS 1e7 3 2>&1 | tail -1 | cut -c1-400
GENERAL=1 timeout -s KILL 300 python tools/prof_run.py ECS 1e7 3 2>&1 | tail -1 | cut -c1-400
