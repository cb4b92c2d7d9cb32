#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_dcs_gpu.py tests/test_ecs_gpu.py tests/test_golden_gpu.py tests/test_chain_gpu.py tests/test_edges_gpu.py -q -m gpu 2>&1 | tail -3
for m in DCS ECS; do timeout -s KILL 300 python tools/prof_run.py $m 1e7 3 2>&1 | tail -1 | cut -c1-200; done
timeout -s KILL 300 python tools/prof_run.py DCS 4e6 3 > gpurun_out/plain_dcs_r2a.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:k_dcs_ -s 1 -c 1 -o gpurun_out/prof_dcs_r2a python tools/prof_run.py DCS 4e6 3 > gpurun_out/ncu_dcs_r2a.log 2>&1
echo "capture rc=$?"
