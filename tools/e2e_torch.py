import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "torch":
    import torch
    torch.cuda.init(); x = torch.zeros(1 << 20, device="cuda"); torch.cuda.synchronize()
    print("torch initialised")
import phasetype_b200 as pb
from phasetype_b200 import synth
wl = synth.config(3, "MHRS", l=10 ** 7)
os.environ["PHT_B200_SEED"] = "1"; os.environ["PHT_B200_QUIET"] = "1"; os.environ["PHT_B200_GPUS"] = "1"
for rep in range(4):
    if rep == 3: os.environ["PHT_B200_TIMING"] = "1"
    t0 = time.perf_counter()
    r = pb.ljma_gibbs(21, 1, 1, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
    print("rep %d: %.3f s" % (rep, time.perf_counter() - t0), flush=True)
