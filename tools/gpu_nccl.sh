#!/bin/bash
NCCL_DEBUG=WARN timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/dbg_nccl.py > gpurun_out/dbg_nccl.log 2>&1; echo "rc=$?"
tail -20 gpurun_out/dbg_nccl.log; cat gpurun_out/dbg_rank0.log gpurun_out/dbg_rank1.log
