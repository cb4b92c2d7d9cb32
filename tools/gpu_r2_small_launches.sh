#!/bin/bash
# per-kernel durations of the once-per-sweep model kernels (assemble, spectral solve, update) at 8 and 32 phases
for cfg in "3 DCS" "5 DCS"; do set -- $cfg
GENERAL=1 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/small_$1.csv python tools/prof_run.py $2 1e5 3 $1 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/small_$1.csv")) if len(r)>5 and r[0].isdigit()]
acc=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0]; v=float(r[-1].replace(",",""))
    a=acc.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
print("config $1:", {k[:28]:(c, round(t/c/1e3,1)) for k,(c,t) in acc.items()}, "(count, mean us)")
PY
done
