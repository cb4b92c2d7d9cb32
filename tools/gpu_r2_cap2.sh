#!/bin/bash
# lane cap above 256 at large shards
for l in 1e7 2e7; do for cap in 256 512 1024; do
  echo -n "l=$l cap=$cap: "; CAP=$cap timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 ns_lane \2 ns_tail \3 ns_replay \4/"
done; done
TRACE=1 timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 6 2>&1 | tail -20
