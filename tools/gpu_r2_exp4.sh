#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_ecs_gpu.py tests/test_dcs_gpu.py tests/test_golden_gpu.py tests/test_chain_gpu.py tests/test_variants_gpu.py tests/test_math.py -q -m gpu -x 2>&1 | tail -3
for m in ECS DCS; do for g in "" 1; do echo -n "$m GENERAL=$g 4e6: "; GENERAL=$g timeout 200 python tools/prof_run.py $m 4e6 3 2>&1 | tail -1 | cut -c1-90; done; done
