"""Tier-1 parity of the CUDA MHRS path against the oracle (and, where the reference build is
present, the unmodified reference C): per-observation B, N and z must be bit-identical."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu


def _engine_paths(R, s, y, cens, mhit, seed, it, cap, world=1, rank=0):
    import phasetype_b200 as pb
    n = s.shape[0]
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    idx = np.arange(rank, y.shape[0], world)
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y[idx], cens[idx], method=1, mhit=mhit, seed=seed,
                    rank=rank, world=world, mhrs_cap=cap, sum_y_global=float(y.sum()))
    eng.set_theta(theta, next_iter=it)
    B, N, z = eng.paths()
    mdl = eng.model()
    cnt = eng.counters()
    eng.close()
    return B, N, z, mdl, cnt, idx


@pytest.mark.parametrize("n,mhit,cap,frac_cens", [(3, 1, 256, 0.0), (4, 2, 4, 0.2), (8, 1, 16, 0.2), (8, 0, 256, 0.5),
                                                 (16, 1, 8, 0.2), (32, 3, 32, 0.1)])
def test_paths_match_oracle(n, mhit, cap, frac_cens):
    rng = np.random.default_rng(100 + n)
    R, s = util.dense_rates(n, rng)
    l = 3000
    y = rng.exponential(1.2, l) + 0.01      # absorption-time scale ~ 1/mean(s), whatever n is
    cens = (rng.uniform(size=l) < frac_cens).astype(np.int32)
    B, N, z, mdl, cnt, _ = _engine_paths(R, s, y, cens, mhit, seed=0xABCDEF12345, it=7, cap=cap)
    Bo, No, zo, co = po.mhrs_paths("oracle", 0xABCDEF12345, 7, y, cens, mdl["S"], mdl["s"], mhit=mhit)
    assert np.array_equal(B, Bo)
    assert np.array_equal(N, No)
    assert np.array_equal(z, zo)            # bit-exact, not a tolerance
    if po.have_ref():
        Br, Nr, zr, _ = po.mhrs_paths("ref", 0xABCDEF12345, 7, y, cens, mdl["S"], mdl["s"], mhit=mhit)
        assert np.array_equal(B, Br) and np.array_equal(N, Nr) and np.array_equal(z, zr)
    if cap < 64:
        assert cnt["deferred"] > 0          # the cooperative tail was exercised


def test_coxian_heavy_tail_and_shards():
    """4-phase Coxian (config 2 shape): a few observations need thousands of attempts; sharded engines
    must reproduce the single-engine result observation by observation (Philox keys are global)."""
    rng = np.random.default_rng(7)
    R, s = util.coxian_rates(4)
    y = util.simulate_pht(R, s, 4000, rng)
    cens = np.zeros(4000, dtype=np.int32)
    B, N, z, mdl, cnt, _ = _engine_paths(R, s, y, cens, 1, seed=99, it=3, cap=64)
    Bo, No, zo, _ = po.mhrs_paths("oracle", 99, 3, y, cens, mdl["S"], mdl["s"], mhit=1)
    assert np.array_equal(B, Bo) and np.array_equal(N, No) and np.array_equal(z, zo)
    assert cnt["deferred"] > 0 and cnt["tail_rounds"] > 0
    for rank in range(2):
        Br, Nr, zr, _, _, idx = _engine_paths(R, s, y, cens, 1, seed=99, it=3, cap=64, world=2, rank=rank)
        assert np.array_equal(Br, Bo[idx]) and np.array_equal(Nr, No[idx]) and np.array_equal(zr, zo[idx])


def test_sweep_stats_equal_sum_of_paths():
    rng = np.random.default_rng(11)
    n = 8
    R, s = util.dense_rates(n, rng)
    T, C, theta = util.general_model(R, s)
    l = 20000
    y = rng.exponential(2.0, l) + 0.01
    cens = (rng.uniform(size=l) < 0.2).astype(np.int32)
    import phasetype_b200 as pb
    eng = pb.Engine(n, T, C, np.full(theta.shape[0], 2.0), np.full(theta.shape[0], 2.0), y, cens, method=1, mhit=1,
                    seed=5, mhrs_cap=32)
    eng.set_theta(theta, next_iter=1)
    Nacc, Bacc, zfix = eng.sweep_stats()
    B, N, z = eng.paths()
    assert np.array_equal(Nacc, N.astype(np.int64).sum(0))
    assert np.array_equal(Bacc, np.bincount(B, minlength=n))
    zf = np.rint(z * 2.0 ** eng.zbits).astype(np.int64).sum(0)
    assert np.array_equal(zfix, zf)
    eng.close()


@pytest.mark.parametrize("cid,l,mhit", [(2, 10 ** 6, 1), (3, 10 ** 6, 1)])
def test_deep_tail_sweep_equals_oracle(cid, l, mhit):
    """The machinery the small cases never reach: default hand-over cap, a million observations, a dozen or more
    cooperative tail rounds with attempts per observation growing to 2^20 and beyond (pool clipping by `found`,
    chunk-major unit numbering, the growth rule).  The packed int64 statistics of the sweep (N | B | z fixed
    point) must equal the CPU restatement's, which walks every observation sequentially."""
    import phasetype_b200 as pb
    from phasetype_b200 import synth
    wl = synth.config(cid, "MHRS", l=l)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, mhit=mhit, seed=0xD33B7A11)
    eng.set_theta(wl.theta, next_iter=4)
    c0 = eng.counters()
    N, B, z = eng.sweep_stats()
    c1 = eng.counters()
    zbits = eng.zbits
    eng.close()
    No, Bo, zo, co = util.oracle_sweep_stats_all_cores(0xD33B7A11, 4, True, mhit, 1, wl.n, wl.T, wl.C, wl.theta, wl.y,
                                                       wl.censored, zbits)
    assert np.array_equal(N, No)
    assert np.array_equal(B, Bo)
    assert np.array_equal(z, zo)
    assert c1["tail_rounds"] - c0["tail_rounds"] >= 6
    assert c1["paths"] - c0["paths"] == l
