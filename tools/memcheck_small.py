"""Tiny run of all three samplers for compute-sanitizer (memcheck): 3000 observations, per-observation paths + 2 sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phasetype_b200 as pb
from phasetype_b200 import synth
for name, code in (("MHRS", 1), ("DCS", 4), ("ECS", 2)):
    wl = synth.config(3, name, l=3000)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=code, seed=3, mhrs_cap=8, use_graph=False)
    eng.set_theta(wl.theta, 1)
    B, N, z = eng.paths()
    out = eng.run(2)
    assert np.isfinite(out).all()
    eng.close()
    print(name, "ok", int(N.sum()))
