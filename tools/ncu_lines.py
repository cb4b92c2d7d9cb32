"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per source line."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []; fname = None; hdr = None
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        try:
            rows.append((fname, int(r[0]), r[1].strip()[:70], int(d["# Samples"]), int(d["Instructions Executed"]),
                         int(d["Thread Instructions Executed"])))
        except Exception: pass
ts = sum(x[3] for x in rows); ti = sum(x[4] for x in rows); tt = sum(x[5] for x in rows)
print("total samples %d, warp instr %d, thread instr %d, avg threads %.1f" % (ts, ti, tt, tt / max(ti, 1)))
rows.sort(key=lambda x: -x[3])
for f, ln, src, s, wi, thi in rows[:top]:
    print("%5.1f%% smp %5.1f%% ins avgthr %4.1f  %s:%d  %s" % (100.0 * s / ts, 100.0 * wi / ti, thi / max(wi, 1), f, ln, src))
