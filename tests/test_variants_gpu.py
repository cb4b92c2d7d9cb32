"""SURVEY 8(f)1 on the GPU: the two sampler variants the reference compiles but never dispatches, behind method bits 8
(LJMA_MHsample_Hobolth, src/Simulate_AbsCTMC_gt_Hobolth_DCS.c:268-355) and 16 (LJMA_MHsample_Aslett,
src/Simulate_AbsCTMC_eq_Aslett_DCS.c:49-143).  Per-observation B, N, z bit-identical to the oracle and to the unmodified
reference symbols (spectral data injected, so all sides consume the same numbers); whole chains through LJMA_Gibbs
against the oracle's driver (the reference's own driver has no branch for them)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu

CODE = {"MHS_HOBOLTH": 8, "MHS_ASLETT": 16}


def _S(R, s):
    n = s.shape[0]; S = R.copy()
    for i in range(n):
        S[i, i] = -(R[i].sum() + s[i])
    return S.ravel(order="F").copy()


@pytest.mark.parametrize("variant", ["MHS_HOBOLTH", "MHS_ASLETT"])
@pytest.mark.parametrize("n,zero_exits,mhit,fc", [(3, False, 1, 0.25), (5, True, 3, 0.25), (8, False, 0, 0.2), (8, False, 2, 0.0),
                                                  (16, False, 1, 0.3), (32, False, 1, 0.2)])
def test_paths_match_oracle_and_reference(variant, n, zero_exits, mhit, fc):
    import phasetype_b200 as pb
    rng = np.random.default_rng(17 * n + mhit)
    R, s = util.dense_rates(n, rng, symmetric=True)
    if zero_exits:
        s[1] = 0.0; s[3] = 0.0
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    S, sv = util.assemble(T, C, theta, n)
    l = 1500 if n <= 8 else (600 if n <= 16 else 200)
    y = util.simulate_pht(R, s, l, rng); cens = (rng.uniform(size=l) < fc).astype(np.int32)
    spec = po.eigen("oracle", S, n)
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=CODE[variant], mhit=mhit, seed=77)
    eng.set_spectral(*spec)
    eng.set_theta(theta, next_iter=3)
    B, N, z = eng.paths()
    stats = eng.sweep_stats(); zbits = eng.zbits
    # a window that does not start at 0 takes the same paths
    B2, N2, z2 = eng.paths(first=l // 3, count=l // 2)
    eng.close()
    Bo, No, zo, _ = po.mh_variant_paths("oracle", variant, 77, 3, y, cens, S, sv, mhit=mhit, spectral=spec)
    assert np.array_equal(B, Bo) and np.array_equal(N, No) and np.array_equal(z, zo)
    if po.have_ref():
        Br, Nr, zr, _ = po.mh_variant_paths("ref", variant, 77, 3, y, cens, S, sv, mhit=mhit, spectral=spec)
        assert np.array_equal(B, Br) and np.array_equal(N, Nr) and np.array_equal(z, zr)
    w = slice(l // 3, l // 3 + l // 2)
    assert np.array_equal(B2, B[w]) and np.array_equal(N2, N[w]) and np.array_equal(z2, z[w])
    Nacc, Bacc, zfix = stats
    assert np.array_equal(Nacc, N.astype(np.int64).sum(0))
    assert np.array_equal(Bacc, np.bincount(B, minlength=n))
    assert np.array_equal(zfix, np.rint(z * 2.0 ** zbits).astype(np.int64).sum(0))


@pytest.mark.parametrize("variant,cid,l", [("MHS_HOBOLTH", 2, 6000), ("MHS_ASLETT", 2, 6000), ("MHS_ASLETT", 4, 1500)])
def test_whole_chain_through_ljma_gibbs(variant, cid, l, monkeypatch):
    import phasetype_b200 as pb
    from phasetype_b200 import synth
    wl = synth.config(cid, "ECS", l=l)
    monkeypatch.setenv("PHT_B200_SEED", "321"); monkeypatch.setenv("PHT_B200_SEED_EXACT", "1"); monkeypatch.setenv("PHT_B200_QUIET", "1")
    monkeypatch.setenv("PHT_B200_GPUS", "1")
    got = pb.ljma_gibbs(5, 2, CODE[variant], wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
    want, _ = po.gibbs(321, 5, 2, CODE[variant], wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    assert np.array_equal(got, want)
