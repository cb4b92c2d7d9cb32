"""Two or more ranks (torchrun, one process per GPU): NCCL all-reduce inside the captured sweep plus the engines'
peer-exchange windows opened across processes through CUDA IPC; the chain must equal the single-process oracle."""
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
from phasetype_b200._lib import PEER_HANDLE_BYTES
L = int(float(os.environ.get("DBG_L", "200000")))
wl = synth.config(3, "MHRS", l=L)
y, c = wl.shard(rank, world)
for graph in ((True,) if os.environ.get("DBG_FAST") else (False, True)):
    e = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, np.ascontiguousarray(y), np.ascontiguousarray(c), method=1, seed=5, device=lr,
                  rank=rank, world=world, use_graph=graph, sum_y_global=float(wl.y.sum()))
    say("engine", graph)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = (C.c_char * 128)()
        assert pb.lib().pht_comm_unique_id(raw) == 0
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0); torch.cuda.synchronize()
    e.comm_init(bytes(buf.cpu().numpy().tobytes()))
    say("comm init done")
    mine = torch.frombuffer(bytearray(e.peer_handle()), dtype=torch.uint8).cuda()
    allh = [torch.zeros(PEER_HANDLE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
    dist.all_gather(allh, mine); torch.cuda.synchronize()
    e.peer_attach(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
    say("peers attached")
    e.set_theta(wl.theta, 1)
    t0 = time.time(); out = e.run(3); dt = time.time() - t0
    cn = e.counters()
    say("run done", dt, out[-1][:3], {k: cn[k] for k in ("tail_rounds", "deferred", "attempts", "jumps", "ns_lane", "ns_tail", "ns_global", "ns_xwait", "global_rounds", "global_items", "ns_replay")})
    e.close()
if rank == 0 and not os.environ.get("DBG_FAST"):
    from oracle import pyoracle as po
    want, _ = po.gibbs(5, 4, 1, 1, wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    say("chain equals single-process oracle:", bool(np.array_equal(out, want[1:])))
dist.barrier(); dist.destroy_process_group()
say("bye")
