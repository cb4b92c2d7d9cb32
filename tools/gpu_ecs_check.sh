#!/bin/bash
timeout -s KILL 600 python -m pytest tests/test_ecs_gpu.py tests/test_chain_gpu.py tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -3
for l in 1e6 4e6; do timeout -s KILL 200 python tools/prof_run.py ECS $l 3 2>&1 | tail -1 | cut -c1-100; done
