"""Per-line warp-instruction share of one source file in an ncu source-page export, in line order.
usage: ncu_file_lines.py report.ncu-rep file.cu [min_pct]"""
import csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]; minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.2
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []; fname = None; hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        try: rows.append((fname, int(r[0]), r[1].strip()[:100], int(d["# Samples"]), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])))
        except Exception: pass
ti = sum(x[4] for x in rows); ts = sum(x[3] for x in rows)
sel = sorted([x for x in rows if x[0] == want], key=lambda x: x[1])
print("file %s: %.1f%% of warp instructions, %.1f%% of samples" % (want, 100.0 * sum(x[4] for x in sel) / ti, 100.0 * sum(x[3] for x in sel) / ts))
for f, ln, src, s, wi, thi in sel:
    if 100.0 * wi / ti >= minp or 100.0 * s / ts >= minp:
        print("%4d ins %5.2f%% smp %5.2f%% thr %4.1f  %s" % (ln, 100.0 * wi / ti, 100.0 * s / ts, thi / max(wi, 1), src))
