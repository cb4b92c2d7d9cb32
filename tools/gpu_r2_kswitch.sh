#!/bin/bash
# usage: gpu_r2_kswitch.sh N K...   (under gpurun --gpus N): strong-scaling sweep time against the K at which the tail goes global
N=$1; shift
for k in "$@"; do
  echo -n "N=$N kswitch=$k: "
  PHT_B200_KSWITCH=$k PHT_BENCH_VERBOSE=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 10 --warmup 3 --no-others --no-cpu --no-e2e > /tmp/ks.json 2> /tmp/ks.err
  python - <<PY
import json
try:
    d=json.loads(open('/tmp/ks.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step'],3), 'ms/sweep, attempts/path', round(d['roofline']['events_per_path']['attempts'],2), 'parity', d.get('chain_parity'))
except Exception as e: print('no json', e)
PY
  grep "rank 0: path kernels" /tmp/ks.err | head -1 | cut -c1-170
done
