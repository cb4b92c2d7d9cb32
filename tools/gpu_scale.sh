#!/bin/bash
# usage: gpu_scale.sh N  (run under gpurun --gpus N)
N=$1
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_n$N.err | grep -v "OMP_NUM\|\*\*\*\*" | tail -5; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches')}, d.get('strong'), d['e2e']['value'], d['roofline']['frac'])
except Exception as e: print('no json', e)
PY
