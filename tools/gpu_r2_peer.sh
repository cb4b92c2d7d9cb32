#!/bin/bash
# usage (under gpurun --gpus N): gpu_r2_peer.sh N -- multi-GPU parity tests (peer all-reduce and NCCL arms), then the bench at N
N=${1:-2}
timeout -s KILL 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -5
bash tools/gpu_scale.sh $N
