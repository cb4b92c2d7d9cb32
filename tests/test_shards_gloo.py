"""The N > 1 host logic on the CPU: two processes (gloo, 127.0.0.1) each reduce their strided shard of the
observations with the CPU restatement, all-reduce the packed int64 statistics block exactly as the engine does
over NCCL, and replay the same parameter draw.  Result must equal the single-process sweep bit for bit."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, method, out):
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as po
    from phasetype_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = synth.config(3, "ECS" if method != 1 else "MHRS", l=1200)
    zb = po.choose_zbits(wl.y.sum())
    theta = wl.theta.copy()
    chain = []
    for it in (1, 2, 3):
        N, B, z, _ = po.sweep_stats(31, it, it == 1, 1, method, wl.n, wl.T, wl.C, theta, wl.y, wl.censored, rank=rank, world=world, zbits=zb)
        block = torch.from_numpy(np.concatenate([N, B, z]))          # the engine's stats block: N | B | z, int64
        dist.all_reduce(block, op=dist.ReduceOp.SUM)
        blk = block.numpy(); n = wl.n
        theta = po.update(31, it, n, wl.nu, wl.zeta, wl.T, wl.C, zb, blk[:n * n], blk[n * n + n:])
        chain.append(theta.copy())
    if rank == 0:
        np.save(out, np.array(chain))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.parametrize("method", [1, 2])
def test_two_rank_sweeps_equal_single_process(method, tmp_path):
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    from phasetype_b200 import synth
    out = str(tmp_path / "chain.npy")
    mp.spawn(_worker, args=(2, _free_port(), method, out), nprocs=2, join=True)
    got = np.load(out)
    wl = synth.config(3, "ECS" if method != 1 else "MHRS", l=1200)
    want, _ = po.gibbs(31, 4, 1, method, wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    assert np.array_equal(got, want[1:])
