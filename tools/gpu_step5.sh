#!/bin/bash
timeout -s KILL 400 python -m pytest tests/test_golden_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --obs 2000000 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_2gpu.log
