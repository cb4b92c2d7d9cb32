"""Tiny driver for ncu: a few sweeps of one method on one GPU (usage: prof_run.py METHOD L SWEEPS [CONFIG])."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phasetype_b200 as pb
from phasetype_b200 import synth
method = sys.argv[1] if len(sys.argv) > 1 else "MHRS"
l = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10 ** 6
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cid = int(sys.argv[4]) if len(sys.argv) > 4 else 3
wl = synth.config(cid, "MHRS" if os.environ.get("GENERAL") else method, l=l)      # GENERAL=1: the unsymmetrised generator (complex spectra)
eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method={"MHRS": 1, "ECS": 2, "DCS": 4}[method],
                mhit=1, seed=1, use_graph=False, mhrs_cap=int(os.environ.get("CAP", "0")))
eng.set_theta(wl.theta, 1)
t = time.time(); out = eng.run(sweeps); dt = time.time() - t
tot, k = eng.last_ms()
print("sweeps", sweeps, "total_ms", tot, "kernel_ms", k, "wall", dt, eng.counters())
if os.environ.get("TRACE"):
    tr = eng.round_trace()
    for r in range(48):
        if tr[r, 4]:
            print("round %2d: search %8.1f us  barrier %7.1f us  advance %7.1f us  P %9.1f  K %10.0f  attempts %11.0f  (%.1f per us)  needed %11.0f  jumps/attempt %.1f" % (r, tr[r, 0] / 1e3 / sweeps, tr[r, 1] / 1e3 / sweeps, tr[r, 2] / 1e3 / sweeps, tr[r, 3] / sweeps, tr[r, 4] / sweeps, tr[r, 5] / sweeps, tr[r, 5] / max(1.0, tr[r, 0] / 1e3), tr[r, 6] / sweeps, tr[r, 7] / max(1.0, tr[r, 5])))
