#!/bin/bash
# build variants of the library (phasetype_b200/build.py --out=... -D...) timed on the same sweeps
for v in libpht_b200.so "$@"; do
  echo -n "== $v: "; PHT_B200_LIB=$PWD/phasetype_b200/$v timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 ns_lane \2 ns_tail \3 ns_replay \4/"
done
