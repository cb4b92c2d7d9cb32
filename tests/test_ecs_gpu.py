"""Tier-1 parity of the CUDA ECS path (exact observations: Aslett-Wilson ECS with per-lane ARMS; censored
observations: Aslett-DCS gt sampler) against the oracle and the unmodified reference C, with the spectral
data injected so both sides consume the same (evals, Q, Q^-1)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu


def _run(R, s, y, cens, seed, it, world=1, rank=0, first=0, count=None):
    import phasetype_b200 as pb
    n = s.shape[0]
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    S, sv = util.assemble(T, C, theta, n)
    spec = po.eigen("oracle", S, n)
    idx = np.arange(rank, y.shape[0], world)
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y[idx], cens[idx], method=2, seed=seed, rank=rank,
                    world=world, sum_y_global=float(y.sum()))
    eng.set_spectral(*spec)
    eng.set_theta(theta, next_iter=it)
    B, N, z = eng.paths(first, count)
    cnt = eng.counters()
    stats = eng.sweep_stats()
    zbits = eng.zbits
    eng.close()
    return B, N, z, S, sv, spec, cnt, idx, stats, zbits


@pytest.mark.parametrize("n,kind,fc", [(3, "dense", 0.0), (3, "dense", 0.3), (4, "coxian", 0.2), (8, "dense", 0.2),
                                       (8, "coxian", 1.0), (16, "dense", 0.2), (32, "dense", 0.2)])
def test_paths_match_oracle(n, kind, fc):
    rng = np.random.default_rng(300 + n)
    R, s = util.dense_rates(n, rng, symmetric=True) if kind == "dense" else util.coxian_rates(n)
    l = 2000 if n <= 16 else 500
    y = rng.exponential(1.2, l) + 0.01
    cens = (rng.uniform(size=l) < fc).astype(np.int32)
    B, N, z, S, sv, spec, cnt, _, stats, zbits = _run(R, s, y, cens, seed=777, it=9)
    Bo, No, zo, co = po.spectral_paths("oracle", "ECS", 777, 9, y, cens, S, sv, spectral=spec)
    assert np.array_equal(B, Bo)
    assert np.array_equal(N, No)
    assert np.array_equal(z, zo)            # bit-exact
    for k in ("jumps", "dens_evals", "env_updates", "arms_calls", "metrop_rejects"):
        assert cnt[k] >= co[k]               # the engine ran paths() and one sweep: at least the oracle's counts
    if po.have_ref():
        Br, Nr, zr, _ = po.spectral_paths("ref", "ECS", 777, 9, y, cens, S, sv, spectral=spec)
        assert np.array_equal(B, Br) and np.array_equal(N, Nr) and np.array_equal(z, zr)
    Nacc, Bacc, zfix = stats
    assert np.array_equal(Nacc, N.astype(np.int64).sum(0))
    assert np.array_equal(Bacc, np.bincount(B, minlength=n))
    assert np.array_equal(zfix, np.rint(z * 2.0 ** zbits).astype(np.int64).sum(0))


def test_subrange_and_shards():
    rng = np.random.default_rng(13)
    n = 8
    R, s = util.dense_rates(n, rng, symmetric=True)
    y = rng.exponential(1.2, 1200) + 0.01
    cens = (rng.uniform(size=1200) < 0.25).astype(np.int32)
    B, N, z, *_ = _run(R, s, y, cens, seed=3, it=4)
    Bs, Ns, zs, *_ = _run(R, s, y, cens, seed=3, it=4, first=100, count=333)
    assert np.array_equal(Bs, B[100:433]) and np.array_equal(Ns, N[100:433]) and np.array_equal(zs, z[100:433])
    for rank in range(2):
        Br, Nr, zr, _, _, _, _, idx, _, _ = _run(R, s, y, cens, seed=3, it=4, world=2, rank=rank)
        assert np.array_equal(Br, B[idx]) and np.array_equal(Nr, N[idx]) and np.array_equal(zr, z[idx])
