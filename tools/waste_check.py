"""How many MHRS attempts does the GPU sweep run beyond what the sequential sampler needs?  (usage: waste_check.py L)
Engine: one sweep at the true parameters, attempts from the device counters (every attempt that ran to its end counts).
Checker: the same sweep through the oracle's sequential rejection loop, split over the host cores."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ProcessPoolExecutor
import phasetype_b200 as pb
from phasetype_b200 import synth
from oracle import pyoracle as po

SEED = 1


def _part(a):
    y, cens, S, s, obs0 = a
    _, _, _, c = po.mhrs_paths("oracle", SEED, 1, y, cens, S, s, mhit=1, obs0=obs0, stride=1, want=False)
    return c["attempts"], c["jumps"]


if __name__ == "__main__":
    l = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 6
    wl = synth.config(3, "MHRS", l=l)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, mhit=1, seed=SEED, use_graph=False)
    eng.set_theta(wl.theta, 1)
    eng.run(1)
    c = eng.counters(); mdl = eng.model()
    cores = os.cpu_count() or 1
    cuts = np.linspace(0, l, cores + 1).astype(int)
    jobs = [(wl.y[a:b], wl.censored[a:b], mdl["S"], mdl["s"], int(a)) for a, b in zip(cuts[:-1], cuts[1:])]
    t = time.time()
    with ProcessPoolExecutor(cores) as ex:
        parts = list(ex.map(_part, jobs))
    att = sum(p[0] for p in parts); jmp = sum(p[1] for p in parts)
    print("l %d: engine attempts %d jumps %d | sequential attempts %d jumps %d | extra attempts %.2f%% (%.2f per path), extra jumps %.2f%% [oracle %.1f s on %d cores]"
          % (l, c["attempts"], c["jumps"], att, jmp, 100.0 * (c["attempts"] - att) / att, (c["attempts"] - att) / l, 100.0 * (c["jumps"] - jmp) / jmp, time.time() - t, cores))
