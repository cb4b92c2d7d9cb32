import csv, sys, collections
path = sys.argv[1]
rows = []; fname = None; hdr = None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        try: rows.append((fname, int(r[0]), int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
        except Exception: pass
g = collections.defaultdict(lambda: [0, 0, 0])
for f, ln, s, wi, ti in rows:
    key = f
    if f == "k_mhrs.cu":
        key = "k_mhrs.cu:jump_step" if 90 <= ln <= 135 else ("k_mhrs.cu:lane" if ln < 270 else "k_mhrs.cu:tail")
    if f == "pht_math.h": key = "pht_log"
    for i, v in enumerate((s, wi, ti)): g[key][i] += v
ts = sum(v[0] for v in g.values()); tw = sum(v[1] for v in g.values())
for k, v in sorted(g.items(), key=lambda kv: -kv[1][1]):
    print("%-32s smp %5.1f%%  warp-instr %5.1f%% (%.3e)  avgthr %4.1f" % (k, 100.0 * v[0] / ts, 100.0 * v[1] / tw, v[1], v[2] / max(v[1], 1)))
