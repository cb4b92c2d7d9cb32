#!/usr/bin/env python
"""bench.py -- Gibbs-sweep throughput of the B200 engine on BASELINE.json's headline workload.

A "step" is one Gibbs sweep (reference src/PHT_MCMC_Aslett.c:268-405) over all observations:
one conditioned latent-path draw per observation + the conjugate parameter update.
Metric: path draws per second (= l x sweeps / time); Gibbs iterations/s is reported beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--method MHRS|ECS|DCS] [--config 3] [--obs L]
  python bench.py --impl reference ...      # the reference's own C on the host cores

Contract details are in the task statement; the JSON line carries `roofline`, `cpu_baseline`,
`e2e`, `clocks`, `gpu_launches`.  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METHOD_CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}
SEED = 0x5048B200


# ----------------------------------------------------------------------------- work model
def work_per_path(method, n, ev):
    """FP64 instruction-equivalents per path, BASELINE.md section 4 (add/mul/fma/compare = 1, div = 8,
    exp = 20, log = 25), from event counts per path measured by the oracle on the benchmark's inputs."""
    if method == "MHRS":
        return ev["jumps"] * (40 + n) + ev["attempts"] * (n + 2)
    if method == "ECS":
        return ev["jumps"] * (2 * n * n + 55 * n + 715) + ev["dens_evals"] * (23 * n + 26) + 550 * ev["env_updates"]
    if method == "DCS":
        return ev["jumps"] * (n * n + 34 * n + 50) + ev["brent_evals"] * (27 * n + 60) + (2 * n * n + 23 * n)
    raise ValueError(method)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device; self.proc = None; self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv"); os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------- reference / oracle on the CPU
def _cpu_worker(args):
    """One process = one single-threaded copy of the reference (it is not re-entrant: globals in src/utility.c:8-12)."""
    kind, method, seed, it, y, cens, S, s, mhit, obs0 = args
    from oracle import pyoracle as po
    t0 = time.perf_counter()
    if method == "MHRS":
        po.mhrs_paths(kind, seed, it, y, cens, S, s, mhit=mhit, obs0=obs0, want=False)
    else:
        po.spectral_paths(kind, method, seed, it, y, cens, S, s, obs0=obs0, want=False)
    return time.perf_counter() - t0


def cpu_sweeps(wl, method, mhit, sample, cores, steps, warmup, kind):
    """Time `steps` sweeps of the CPU implementation over the first `sample` observations, split over
    `cores` independent processes.  Returns (paths_per_s, ms_per_step)."""
    import multiprocessing as mp
    from oracle import pyoracle as po
    S, s = model_matrices(wl)
    y = wl.y[:sample]; c = wl.censored[:sample]
    bounds = np.linspace(0, sample, cores + 1).astype(int)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for k in range(warmup + steps):
            jobs = [(kind, method, SEED, 1 + k, y[a:b], c[a:b], S, s, mhit, int(a)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if k >= warmup:
                times.append(dt)
    tot = float(np.sum(times))
    return sample * steps / tot, 1e3 * tot / steps


def model_matrices(wl):
    n = wl.n
    S = wl.R.copy()
    for i in range(n):
        S[i, i] = -(wl.R[i].sum() + wl.s[i])
    return np.asfortranarray(S).ravel(order="F").copy(), wl.s.copy()


def ref_kind():
    from oracle import pyoracle as po
    po.build()
    return "ref" if po.have_ref() else "oracle"


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="MHRS", choices=["MHRS", "ECS", "DCS"])
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--obs", dest="l", type=int, default=None, help="override the number of observations")
    ap.add_argument("--mhit", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from phasetype_b200 import synth

    if args.impl == "reference":
        if rank != 0:
            return 0
        wl = synth.config(args.config, args.method, l=min(args.l or 10 ** 7, 2 * 10 ** 6))
        full_l = args.l or {1: 100, 2: 10 ** 6, 3: 10 ** 7, 4: 10 ** 7, 5: 10 ** 8}[args.config]
        cores = os.cpu_count() or 1
        kind = ref_kind()
        sample = args.cpu_sample or min(wl.l, 20000 * cores)
        pps, ms = cpu_sweeps(wl, args.method, args.mhit, sample, cores, args.steps, args.warmup, kind)
        line = {"impl": "reference", "metric": "path_draws_per_sec", "value": pps, "unit": "paths/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "gibbs_iters_per_sec_extrapolated": pps / full_l,
                "config": {"workload": wl.name, "method": args.method, "mhit": args.mhit, "phases": wl.n,
                           "observations": full_l, "sample_per_step": sample},
                "cpu_baseline": {"value": pps, "unit": "paths/s", "cores": cores,
                                 "kind": "reference" if kind == "ref" else "port",
                                 "sample": "%d observations per step, %d single-threaded processes" % (sample, cores)},
                "e2e": {"value": pps, "unit": "paths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import phasetype_b200 as pb
    dist = None; torch = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = synth.config(args.config, args.method, l=args.l)
    l = wl.l
    y_loc, c_loc = wl.shard(rank, world)
    y_loc = np.ascontiguousarray(y_loc); c_loc = np.ascontiguousarray(c_loc)
    sum_y = float(wl.y.sum())
    code = METHOD_CODE[args.method]

    def make_engine(graph):
        e = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, y_loc, c_loc, method=code, mhit=args.mhit, seed=SEED,
                      device=local_rank, rank=rank, world=world, use_graph=graph, sum_y_global=sum_y)
        if world > 1:
            buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                import ctypes as C
                raw = (C.c_char * 128)()
                if pb.lib().pht_comm_unique_id(raw) != 0:
                    raise RuntimeError(pb.lib().pht_last_error().decode())
                buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
            dist.broadcast(buf, 0)
            e.comm_init(bytes(buf.cpu().numpy().tobytes()))
        e.set_theta(wl.theta, next_iter=1)
        return e

    def barrier():
        if world > 1:
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = make_engine(True)
    eng.run(args.warmup)
    c0 = eng.counters()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    eng.enqueue(args.steps); eng.sync()
    t_wall = time.perf_counter() - t_wall0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms, _ = eng.last_ms()
    total_ms = max_over_ranks(total_ms)
    c1 = eng.counters()
    launches = c1["launches"] - c0["launches"]
    ev_gpu = {k: (c1[k] - c0[k]) / float(args.steps * max(1, y_loc.shape[0])) for k in ("attempts", "jumps", "deferred")}
    tail_rounds = (c1["tail_rounds"] - c0["tail_rounds"]) / float(args.steps)
    value = l * args.steps / (total_ms * 1e-3)
    eng.close()

    # ---- instrumented pass: average duration of the path kernel alone (CUDA events around each launch)
    eng2 = make_engine(False)
    eng2.run(args.warmup)
    barrier()
    eng2.enqueue(min(args.steps, 32)); eng2.sync()
    tot2, kern_ms = eng2.last_ms()
    kern_ms = max_over_ranks(kern_ms)
    share = kern_ms * min(args.steps, 32) / max(tot2, 1e-9)
    eng2.close()

    line = None
    if rank == 0:
        fma_rate = pb.fp64_fma_rate(local_rank)
        line = {"metric": "path_draws_per_sec", "value": value, "unit": "paths/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "gibbs_iters_per_sec": args.steps / (total_ms * 1e-3),
                "config": {"workload": wl.name, "method": args.method, "mhit": args.mhit, "phases": wl.n,
                           "parameters": wl.m, "observations": l, "sharding": "observation i -> rank i mod %d" % world,
                           "l2_note": "inputs (%.0f MB per GPU) are streamed once per sweep; each sweep re-reads them after ~%.1f ms of unrelated work, and at l >= 1e7 they exceed no cache assumption because the kernel is FP64-issue bound (12 B per path)" % (9e-6 * y_loc.shape[0], total_ms / args.steps)},
                "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * t_wall / args.steps,
                "device_events_per_path": ev_gpu, "tail_rounds_per_sweep": tail_rounds,
                "clocks": clocks}
    # ---- roofline (rank 0): algorithmic work from oracle event counts on the benchmark's own inputs
    if rank == 0:
        from oracle import pyoracle as po
        po.build()
        S, s = model_matrices(wl)
        ns = min(l, 20000)
        if args.method == "MHRS":
            _, _, _, cnt = po.mhrs_paths("oracle", SEED, 1, wl.y[:ns], wl.censored[:ns], S, s, mhit=args.mhit, want=False)
        else:
            cnt = po.spectral_paths("oracle", args.method, SEED, 1, wl.y[:ns], wl.censored[:ns], S, s, want=False)[-1]
        ev = {k: v / float(ns) for k, v in cnt.items()}
        W = work_per_path(args.method, wl.n, ev)
        per_launch = W * y_loc.shape[0]
        achieved = per_launch / (kern_ms * 1e-3)
        line["roofline"] = {"bound": "fp64_issue", "achieved": achieved / 1e9, "peak": fma_rate / 1e9, "unit": "G FP64-instr-equiv/s",
                            "frac": achieved / fma_rate, "traffic": None,
                            "peak_source": "measured in this run: dependent FP64 FMA chains on all SMs (pht_fp64_fma_rate)",
                            "kernel": "k_%s_sweep" % args.method.lower(), "kernel_ms": kern_ms, "kernel_share_of_step": share,
                            "work_per_path": W, "events_per_path": ev,
                            "hbm": {"algorithmic_bytes_per_launch": 9 * y_loc.shape[0],
                                    "achieved_GBs": 9e-9 * y_loc.shape[0] / (kern_ms * 1e-3), "peak_GBs": _peak_hbm()}}

    # ---- end to end through the drop-in routine with host buffers (upload, sweeps, download)
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            os.environ["PHT_B200_SEED"] = str(SEED); os.environ["PHT_B200_QUIET"] = "1"
            os.environ["PHT_B200_DEVICE"] = str(local_rank)
            res = pb.ljma_gibbs(args.steps + 1, args.mhit, code, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y,
                                wl.censored, wl.theta, silent=True)
            assert np.isfinite(res).all() and (res[1:] > 0).all()
        else:
            e3 = make_engine(True)
            res = e3.run(args.steps)
            e3.close()
        dt = time.perf_counter() - t0
        if world > 1:
            dt = max_over_ranks(dt)
        e2e = {"value": l * args.steps / dt, "unit": "paths/s", "h2d_bytes_per_step": int(12 * l / args.steps),
               "d2h_bytes_per_step": int(8 * wl.m),
               "note": "LJMA_Gibbs(it=%d) on host vectors: engine creation, upload of y/censored, %d sweeps, download of res" % (args.steps + 1, args.steps)
               if world == 1 else "engine create from host shards + sweeps + result download on every rank"}
    if rank == 0:
        line["e2e"] = e2e
        if not args.no_cpu:
            cores = 1
            kind = "ref" if po.have_ref() else "oracle"
            sample = args.cpu_sample or min(l, 100000)
            pps, ms = cpu_sweeps(wl, args.method, args.mhit, sample, cores, 3, 1, kind)
            line["cpu_baseline"] = {"value": pps, "unit": "paths/s", "cores": cores,
                                    "kind": "reference" if kind == "ref" else "port",
                                    "sample": "first %d observations, 3 sweeps, 1 process (the reference is single-threaded)" % sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


def _peak_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    sys.exit(main())
