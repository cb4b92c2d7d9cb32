#!/bin/bash
# the two MH variants (method bits 8 and 16) against the oracle and the reference symbols, then the live DCS / ECS parity tests
timeout -s KILL 900 python -m pytest tests/test_variants_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout -s KILL 900 python -m pytest tests/test_dcs_gpu.py tests/test_ecs_gpu.py tests/test_golden_gpu.py tests/test_chain_gpu.py -q -m gpu 2>&1 | tail -3
