"""Summarise an .ncu-rep for profiles/: key counters per kernel launch + the hottest source lines.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_lines] > profiles/<name>.md
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per warp instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (fp32/imad) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "thread DFMA"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "thread DMUL"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "thread DADD"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed"),
]
STALLS = "smsp__average_warp"


def ncu(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"] + list(extra), capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_section(rep):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append("### %s  (launch id %s)\n" % (d.get("Kernel Name", "?"), d.get("ID", "?")))
        out.append("| counter | value |\n|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                out.append("| %s (`%s`) | %s %s |" % (label, k, d[k], units[hdr.index(k)]))
        stalls = []
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d.get(k):
                try:
                    stalls.append((float(d[k]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            out.append("\nwarp stall reasons (warps stalled per issue-active cycle): " +
                       ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))
        out.append("")
    return "\n".join(out)


def source_section(rep, top):
    rows = ncu(rep, "source", ["--print-source", "cuda,sass"])
    recs = []; fname = None; hdr = None; kern = None
    per_kernel = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "Kernel Name":
            kern = r[1]; continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
            d = dict(zip(hdr[4:], r[4:]))
            try:
                recs.append((fname, int(r[0]), r[1].strip()[:90], int(d["# Samples"]), int(d["Instructions Executed"]),
                             int(d["Thread Instructions Executed"])))
            except Exception:
                pass
    if not recs:
        return "(no source page in this report)\n"
    ts = sum(x[3] for x in recs) or 1; ti = sum(x[4] for x in recs) or 1; tt = sum(x[5] for x in recs)
    out = ["all profiled launches together: %d samples, %.3e warp instructions, %.1f active threads per instruction\n" % (ts, ti, tt / ti),
           "| samples % | warp instr % | active threads | line | source |\n|---|---|---|---|---|"]
    recs.sort(key=lambda x: -x[3])
    for f, ln, src, s, wi, thi in recs[:top]:
        out.append("| %.1f | %.1f | %.1f | %s:%d | `%s` |" % (100.0 * s / ts, 100.0 * wi / ti, thi / max(wi, 1), f, ln, src.replace("|", "\\|")))
    # per-file totals
    tot = {}
    for f, ln, src, s, wi, thi in recs:
        a = tot.setdefault(f, [0, 0, 0]); a[0] += s; a[1] += wi; a[2] += thi
    out.append("\n| file | samples % | warp instr % | active threads |\n|---|---|---|---|")
    for f, a in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        out.append("| %s | %.1f | %.1f | %.1f |" % (f, 100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1)))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    print("# ncu summary of `%s`\n" % rep.split("/")[-1])
    print("Captured with `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays:\n"
          "use shares and ratios, not absolute times).\n")
    print("## Launches\n")
    print(raw_section(rep))
    print("## Hottest source lines\n")
    print(source_section(rep, top))
