"""Tier-1 parity of the CUDA DCS path (Aslett-Hobolth) against the oracle and, where present, the
unmodified reference C: per-observation B, N, z bit-identical when both sides consume the same
spectral data (evals, Q, Q^-1 from LAPACK, injected through pht_engine_set_spectral)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu


def _run(R, s, y, seed, it, world=1, rank=0):
    import phasetype_b200 as pb
    n = s.shape[0]
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    S, sv = util.assemble(T, C, theta, n)
    spec = po.eigen("oracle", S, n)
    cens = np.zeros(y.shape[0], dtype=np.int32)
    idx = np.arange(rank, y.shape[0], world)
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y[idx], cens[idx], method=4, seed=seed, rank=rank,
                    world=world, sum_y_global=float(y.sum()))
    eng.set_spectral(*spec)
    eng.set_theta(theta, next_iter=it)
    B, N, z = eng.paths()
    mdl = eng.model()
    assert np.array_equal(mdl["S"], S) and np.array_equal(mdl["s"], sv)
    cnt = eng.counters()
    stats = eng.sweep_stats()
    zbits = eng.zbits
    eng.close()
    return B, N, z, S, sv, spec, cnt, idx, stats, zbits


@pytest.mark.parametrize("n,kind", [(3, "dense"), (4, "coxian"), (8, "dense"), (16, "dense"), (32, "dense")])
def test_paths_match_oracle(n, kind):
    rng = np.random.default_rng(200 + n)
    R, s = util.dense_rates(n, rng, symmetric=True) if kind == "dense" else util.coxian_rates(n)
    l = 2000 if n <= 16 else 600
    y = rng.exponential(1.2, l) + 0.01
    B, N, z, S, sv, spec, cnt, _, stats, zbits = _run(R, s, y, seed=4242, it=5)
    cens = np.zeros(l, dtype=np.int32)
    Bo, No, zo, co = po.spectral_paths("oracle", "DCS", 4242, 5, y, cens, S, sv, spectral=spec)
    assert np.array_equal(B, Bo)
    assert np.array_equal(N, No)
    assert np.array_equal(z, zo)            # bit-exact
    assert cnt["jumps"] >= co["jumps"] and cnt["brent_evals"] >= co["brent_evals"]
    if po.have_ref():
        Br, Nr, zr, _ = po.spectral_paths("ref", "DCS", 4242, 5, y, cens, S, sv, spectral=spec)
        assert np.array_equal(B, Br) and np.array_equal(N, Nr) and np.array_equal(z, zr)
    Nacc, Bacc, zfix = stats
    assert np.array_equal(Nacc, N.astype(np.int64).sum(0))
    assert np.array_equal(Bacc, np.bincount(B, minlength=n))
    assert np.array_equal(zfix, np.rint(z * 2.0 ** zbits).astype(np.int64).sum(0))


def test_shards_reproduce_global_paths():
    rng = np.random.default_rng(9)
    n = 8
    R, s = util.dense_rates(n, rng, symmetric=True)
    y = rng.exponential(1.2, 1500) + 0.01
    B, N, z, *_ = _run(R, s, y, seed=1, it=2)
    for rank in range(3):
        Br, Nr, zr, _, _, _, _, idx, _, _ = _run(R, s, y, seed=1, it=2, world=3, rank=rank)
        assert np.array_equal(Br, B[idx]) and np.array_equal(Nr, N[idx]) and np.array_equal(zr, z[idx])
