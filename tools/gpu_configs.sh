#!/bin/bash
# throughput of the other BASELINE configs (shape checks at 1e6 observations; usage: prof_run.py METHOD L SWEEPS CONFIG)
for c in "MHRS 1e6 3 2" "ECS 1e6 3 2" "DCS 1e6 3 2" "MHRS 1e6 3 4" "ECS 1e6 3 4" "DCS 1e6 3 4" "MHRS 1e6 3 5" "DCS 5e5 3 5" "ECS 5e5 3 5"; do
  set -- $c; echo "== config $4 $1 l=$2"; timeout -s KILL 300 python tools/prof_run.py $1 $2 $3 $4 2>&1 | tail -1 | cut -c1-400
done
