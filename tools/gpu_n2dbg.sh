#!/bin/bash
# usage: gpu_n2dbg.sh N : a few 1e7-observation sweeps sharded over N ranks with the tail's phase timers, for several k_switch values
N=${1:-2}
for ks in 32768 262144 4194304; do
  echo "== k_switch $ks"
  PHT_B200_KSWITCH=$ks DBG_FAST=1 DBG_L=1e7 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/dbg_nccl.py > gpurun_out/dbg_nccl.log 2>&1; echo "rc=$?"
  for r in 0 $((N-1)); do grep "run done" gpurun_out/dbg_rank$r.log | cut -c1-420; done
done
