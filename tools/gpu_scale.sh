#!/bin/bash
# usage: gpu_scale.sh N [extra bench args]  (run under gpurun --gpus N)
N=$1; shift
if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py"; fi
PHT_BENCH_VERBOSE=1 timeout -s KILL 1200 $CMD --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -12 | cut -c1-330; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches','chain_parity')}, 'weak', d.get('weak'), 'e2e', d['e2e'] and d['e2e']['value'], 'frac', d['roofline']['frac'])
except Exception as e: print('no json', e)
PY
