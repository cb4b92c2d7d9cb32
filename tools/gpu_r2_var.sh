#!/bin/bash
# build variants of the library (/tmp/mkvar.sh NAME FILE -D...) timed on the same MHRS sweeps
for v in libpht_b200.so "$@"; do
  for l in 1.25e6 1e7; do echo -n "== $v l=$l: "; PHT_B200_LIB=$PWD/phasetype_b200/$v timeout -s KILL 200 python tools/prof_run.py MHRS $l 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'attempts': ([0-9]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 attempts \2 ns_lane \3 ns_tail \4 ns_replay \5/"; done
done
