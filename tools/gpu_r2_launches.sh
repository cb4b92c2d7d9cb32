#!/bin/bash
# ncu launch lists (gpu__time_duration per launch): the bench command, and a small-shard run (one GPU's share of the 8-GPU job)
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 5 --warmup 3 --no-others --no-cpu --no-e2e > gpurun_out/ncu_launches_r2.log 2>&1
echo "bench launches rc=$?"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_small_launches.csv python tools/prof_run.py MHRS 1.25e6 6 > gpurun_out/ncu_small_r2.log 2>&1
echo "small launches rc=$?"
python - <<'PY'
import csv,collections
for f in ("gpurun_out/r2_bench_launches.csv","gpurun_out/r2_small_launches.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and r[0].isdigit()]
    acc=collections.OrderedDict()
    for r in rows:
        name=r[4].split("(")[0]; v=float(r[-1].replace(",","")); u=r[-2]
        a=acc.setdefault(name,[0,0.0,[]]); a[0]+=1; a[1]+=v; a[2].append(v)
    print(f)
    for k,(c,t,l) in acc.items(): print("  %-40s n=%3d total %12.1f  last %s %s"%(k[:40],c,t,[round(x) for x in l[-3:]],u))
PY
