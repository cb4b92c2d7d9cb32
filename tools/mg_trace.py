"""Several GPUs driven from ONE process (an engine and a host thread per device, exchange windows attached by direct
peer pointers, as LJMA_Gibbs does it): a few MHRS sweeps of the sharded workload with rank 0's per-round tail trace.
usage: mg_trace.py N L SWEEPS"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import phasetype_b200 as pb
from phasetype_b200 import synth
N = int(sys.argv[1]); L = int(float(sys.argv[2])); sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
wl = synth.config(3, "MHRS", l=L)
engs = []
for r in range(N):
    y, c = wl.shard(r, N)
    engs.append(pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, np.ascontiguousarray(y), np.ascontiguousarray(c), method=1, seed=5, device=r,
                          rank=r, world=N, use_graph=False, sum_y_global=float(wl.y.sum())))
if N > 1:
    handles = b"".join(e.peer_handle() for e in engs)
    for e in engs:
        e.peer_attach(handles)
for e in engs:
    e.set_theta(wl.theta, 1)
outs = [None] * N
def work(r):
    outs[r] = engs[r].run(sweeps)
th = [threading.Thread(target=work, args=(r,)) for r in range(N)]
t0 = time.time()
for t in th: t.start()
for t in th: t.join()
dt = time.time() - t0
assert all(np.array_equal(o, outs[0]) for o in outs)
tot, k = engs[0].last_ms(); cn = engs[0].counters()
print("N", N, "L", L, "sweeps", sweeps, "wall %.3f s" % dt, "kernel_ms %.3f" % k, {q: cn[q] for q in ("deferred", "tail_rounds", "global_rounds", "global_items", "attempts")})
print("per sweep ms: lanes %.3f local tail %.3f global tail %.3f (waiting for peers %.3f) replay %.3f" % tuple(cn[q] * 1e-6 / sweeps for q in ("ns_lane", "ns_tail", "ns_global", "ns_xwait", "ns_replay")))
tr = engs[0].round_trace()
for r in range(48):
    if tr[r, 4]:
        print("round %2d: search %8.1f us  barrier %7.1f us  advance %7.1f us  P %9.1f  K %10.0f  attempts(this rank) %11.0f  (%.1f per us)" % (r, tr[r, 0] / 1e3 / sweeps, tr[r, 1] / 1e3 / sweeps, tr[r, 2] / 1e3 / sweeps, tr[r, 3] / sweeps, tr[r, 4] / sweeps, tr[r, 5] / sweeps, tr[r, 5] / max(1.0, tr[r, 0] / 1e3)))
for e in engs:
    e.close()
