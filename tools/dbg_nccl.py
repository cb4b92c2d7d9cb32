import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
log = open("gpurun_out/dbg_rank%d.log" % rank, "w")
def say(*a):
    print(time.strftime("%H:%M:%S"), *a, file=log, flush=True)
import torch, torch.distributed as dist
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
say("pg up")
import phasetype_b200 as pb, ctypes as C
from phasetype_b200 import synth
wl = synth.config(3, "MHRS", l=200000)
y, c = wl.shard(rank, world)
for graph in (False, True):
    e = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, np.ascontiguousarray(y), np.ascontiguousarray(c), method=1, seed=5, device=lr,
                  rank=rank, world=world, use_graph=graph, sum_y_global=float(wl.y.sum()))
    say("engine", graph)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_char * 128)()
        assert pb.lib().pht_comm_unique_id(raw) == 0
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0); torch.cuda.synchronize()
    say("id broadcast")
    e.comm_init(bytes(buf.cpu().numpy().tobytes()))
    say("comm init done")
    e.set_theta(wl.theta, 1)
    out = e.run(3)
    say("run done", out[-1][:3])
    e.close()
if rank == 0:
    from oracle import pyoracle as po
    want, _ = po.gibbs(5, 4, 1, 1, wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    say("chain equals single-process oracle:", bool(np.array_equal(out, want[1:])))
dist.barrier(); dist.destroy_process_group()
say("bye")
