#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_ecs_gpu.py tests/test_variants_gpu.py tests/test_complex_gpu.py tests/test_golden_gpu.py tests/test_chain_gpu.py tests/test_tiers_gpu.py tests/test_edges_gpu.py -q -m gpu -x 2>&1 | tail -3
for g in 1 ""; do echo -n "ECS GENERAL=$g 4e6: "; GENERAL=$g timeout 200 python tools/prof_run.py ECS 4e6 3 2>&1 | tail -1 | cut -c1-90; done
echo -n "ECS general 1e7: "; GENERAL=1 timeout 200 python tools/prof_run.py ECS 1e7 3 2>&1 | tail -1 | cut -c1-90
