#!/bin/bash
# lane cap: the automatic rule against fixed values, at the shard sizes of the 1/2/4/8-GPU runs of the 1e7-observation job
for L in 1.25e6 2.5e6 5e6 1e7; do
  for C in 0 256 64 32 16; do
    echo -n "L=$L CAP=$C: "; CAP=$C timeout -s KILL 200 python tools/prof_run.py MHRS $L 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'deferred': ([0-9]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 deferred \2 ns_lane \3 ns_tail \4 ns_replay \5/"
  done
done
