#!/bin/bash
# lane cap at one GPU's share of the 8- and 4- and 2-GPU runs
for l in 1.25e6 2.5e6 5e6; do for cap in 32 64 128 256 512; do
  echo -n "l=$l cap=$cap: "; CAP=$cap timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'attempts': ([0-9]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 attempts \2 ns_lane \3 ns_tail \4 ns_replay \5/"
done; done
