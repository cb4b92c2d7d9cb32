#!/bin/bash
# occupancy variants of the ECS / DCS kernels on the general C3 (4e6 observations, 3 sweeps), plus the MHRS default
t() { PHT_B200_LIB=$PWD/phasetype_b200/$1 GENERAL=1 timeout -s KILL 200 python tools/prof_run.py $2 4e6 3 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*/\1/"; }
for v in libpht_b200.so libpht_e20.so libpht_e12.so; do echo "ECS $v: $(t $v ECS)"; done
for v in libpht_b200.so libpht_d24.so libpht_d16.so libpht_du4.so; do echo "DCS $v: $(t $v DCS)"; done
echo -n "MHRS 1e7: "; timeout -s KILL 200 python tools/prof_run.py MHRS 1e7 6 2>&1 | tail -1 | sed -E "s/.*kernel_ms ([0-9.]+).*'ns_lane': ([0-9]+), 'ns_tail': ([0-9]+), 'ns_replay': ([0-9]+).*/kernel_ms \1 ns_lane \2 ns_tail \3 ns_replay \4/"
