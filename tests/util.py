"""Shared helpers for the tests: synthetic phase-type models in LJMA_Gibbs argument form."""
import numpy as np


def dense_rates(n, rng, symmetric=False, lo=0.2, hi=1.2):
    """Off-diagonal rates R (n x n, zero diagonal) and exit rates s."""
    R = rng.uniform(lo, hi, (n, n))
    if symmetric:
        R = (R + R.T) / 2
    np.fill_diagonal(R, 0.0)
    s = rng.uniform(lo, hi, n)
    return R, s


def coxian_rates(n, early_exit=0.3, last_exit=1.5):
    """forward rates 1.0 + 0.37 i, early-exit rate, last exit."""
    R = np.zeros((n, n)); s = np.full(n, early_exit)
    for i in range(n - 1):
        R[i, i + 1] = 1.0 + 0.37 * i
    s[n - 1] = last_exit
    return R, s


def general_model(R, s):
    """Every non-zero rate is its own parameter.  Returns (T, C, theta): T, C column-major (n+1)^2
    in the reference's encoding (R/phtMCMC2.R:58-63) with parameters numbered column-major."""
    n = s.shape[0]
    T = np.zeros((n + 1, n + 1), dtype=np.int32)
    theta = []
    for j in range(n + 1):
        for i in range(n):
            v = s[i] if j == n else R[i, j]
            if v != 0.0 and i != j:
                theta.append(v); T[i, j] = len(theta)
    C = np.ones((n + 1, n + 1))
    return T.ravel(order="F").copy(), C.ravel(order="F").copy(), np.array(theta)


def simulate_pht(R, s, size, rng):
    """Absorption times of the CTMC started in state 0 (plain numpy, test data only)."""
    n = s.shape[0]
    rate = R.sum(1) + s
    P = np.concatenate([R, s[:, None]], axis=1) / rate[:, None]
    out = np.zeros(size)
    for k in range(size):
        j = 0; t = 0.0
        while j < n:
            t += rng.exponential(1.0 / rate[j])
            j = rng.choice(n + 1, p=P[j])
        out[k] = t
    return out


def assemble(T, C, theta, n, first=True):
    """S (column-major flat) and s from the parameter vector exactly as the engine's k_assemble / the reference
    do it: cell = theta * C, diagonal = -(row sum) accumulated over ascending columns for the start values
    (src/PHT_MCMC_Aslett.c:215,229) and over descending columns afterwards (:389-393)."""
    n1 = n + 1
    T = np.asarray(T).reshape(n1, n1, order="F"); C = np.asarray(C, dtype=np.float64).reshape(n1, n1, order="F")
    TT = np.zeros((n1, n1))
    for i in range(n1):
        for j in range(n1):
            if T[i, j] != 0:
                TT[i, j] = theta[T[i, j] - 1] * C[i, j]
    for i in range(n):
        acc = 0.0
        cols = range(n1) if first else range(n, -1, -1)
        for j in cols:
            if T[i, j] != 0:
                acc -= TT[i, j]
        TT[i, i] = acc
    S = TT[:n, :n].ravel(order="F").copy()
    s = TT[:n, n].copy()
    return S, s


def _oracle_shard_stats(args):
    from oracle import pyoracle as po
    seed, it, first, mhit, method, n, T, C, theta, y, cens, rank, world, zbits = args
    N, B, z, cnt = po.sweep_stats(seed, it, first, mhit, method, n, T, C, theta, y, cens, rank=rank, world=world, zbits=zbits)
    return N, B, z, cnt


def oracle_sweep_stats_all_cores(seed, it, first, mhit, method, n, T, C, theta, y, cens, zbits):
    """One sweep's packed statistics from the CPU restatement, the observations split over every host core
    (int64 sums: the split cannot change the totals)."""
    import multiprocessing as mp
    import os
    w = max(1, os.cpu_count() or 1)
    jobs = [(seed, it, first, mhit, method, n, T, C, theta, y, cens, r, w, zbits) for r in range(w)]
    with mp.get_context("fork").Pool(w) as pool:
        out = pool.map(_oracle_shard_stats, jobs)
    N = sum(o[0] for o in out); B = sum(o[1] for o in out); z = sum(o[2] for o in out)
    cnt = {k: sum(o[3][k] for o in out) for k in out[0][3]}
    return N, B, z, cnt
