#!/bin/bash
# usage: gpu_nccl.sh N   (under gpurun --gpus N): torchrun path with IPC peer windows, then the in-process multi-GPU tests
N=${1:-2}
NCCL_DEBUG=WARN timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg_nccl.py > gpurun_out/dbg_nccl.log 2>&1; echo "torchrun rc=$?"
tail -5 gpurun_out/dbg_nccl.log | cut -c1-300; for r in $(seq 0 $((N-1))); do tail -4 gpurun_out/dbg_rank$r.log | cut -c1-400; done
timeout -s KILL 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x 2>&1 | tail -15
