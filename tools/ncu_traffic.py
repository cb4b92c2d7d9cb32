"""Record the DRAM traffic per launch of a path kernel from an `ncu --set full` capture into profiles/traffic.json,
where bench.py picks it up as roofline.traffic.
usage: python tools/ncu_traffic.py report.ncu-rep METHOD L_LOCAL [TAG]     (TAG, e.g. "general": key METHOD:TAG:L_LOCAL)"""
import csv, io, json, os, subprocess, sys
rep, method, l_local = sys.argv[1], sys.argv[2], int(float(sys.argv[3]))
tag = sys.argv[4] if len(sys.argv) > 4 else None
key = "%s:%s:%d" % (method, tag, l_local) if tag else "%s:%d" % (method, l_local)
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]
def to_bytes(v, u):
    v = float(v); u = u.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
per_kernel = {}; missing = []; winst = {}; lanes = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0]
    try:
        rd = to_bytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
    except ValueError:
        continue
    if rd != rd or wr != wr:          # ncu could not collect the counters for this launch
        missing.append(name); continue
    per_kernel.setdefault(name, []).append(rd + wr)
    try:
        winst.setdefault(name, []).append(float(d["smsp__inst_executed.sum"]))
        lanes.setdefault(name, []).append(float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]))
    except (KeyError, ValueError):
        pass
total = sum(sum(v) / len(v) for v in per_kernel.values())      # one launch of each kernel of the method per sweep
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
t = json.load(open(path)) if os.path.exists(path) else {}
t[key] = {"dram_bytes_per_launch": total, "kernels": {k: sum(v) / len(v) for k, v in per_kernel.items()},
                                   "source": os.path.basename(rep), "algorithmic_bytes": 9 * l_local,
                                   "not_collected": sorted(set(missing)),
                                   "warp_instructions_per_launch": sum(sum(v) / len(v) for v in winst.values()),
                                   "active_lanes_per_instruction": {k: sum(v) / len(v) for k, v in lanes.items()}}
json.dump(t, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(t[key]))
