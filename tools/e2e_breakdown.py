"""Where the time of one LJMA_Gibbs call goes at 1e7 observations (engine creation / sweeps / destruction)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phasetype_b200 as pb
from phasetype_b200 import synth
wl = synth.config(3, "MHRS", l=10 ** 7)
os.environ["PHT_B200_SEED"] = "1"; os.environ["PHT_B200_QUIET"] = "1"
for rep in range(3):
    t0 = time.perf_counter()
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, seed=1)
    t1 = time.perf_counter()
    eng.set_theta(wl.theta, 1)
    out = eng.run(20)
    t2 = time.perf_counter()
    eng.close()
    t3 = time.perf_counter()
    r = pb.ljma_gibbs(21, 1, 1, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
    t4 = time.perf_counter()
    print("rep %d: create %.3f s, 20 sweeps %.3f s, destroy %.3f s | LJMA_Gibbs(it=21) %.3f s" % (rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3))
os.environ["PHT_B200_TIMING"] = "1"
t0 = time.perf_counter()
r = pb.ljma_gibbs(21, 1, 1, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
print("timed call: %.3f s" % (time.perf_counter() - t0))
