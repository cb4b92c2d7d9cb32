"""Stage times of one multi-GPU LJMA_Gibbs call (usage: e2e_multi.py GPUS [SWEEPS])."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phasetype_b200 as pb
from phasetype_b200 import synth
g = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = synth.config(3, "MHRS", l=10 ** 7)
os.environ["PHT_B200_SEED"] = "1"; os.environ["PHT_B200_QUIET"] = "1"; os.environ["PHT_B200_GPUS"] = str(g)
for rep in range(4):
    os.environ["PHT_B200_TIMING"] = "1"
    t0 = time.perf_counter()
    r = pb.ljma_gibbs(sweeps + 1, 1, 1, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
    print("rep %d: LJMA_Gibbs(it=%d) on %d GPUs %.3f s" % (rep, sweeps + 1, g, time.perf_counter() - t0), flush=True)
