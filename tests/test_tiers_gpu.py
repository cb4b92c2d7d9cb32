"""Tiers 2 and 3 of the correctness contract measured directly on the CUDA path (tier 1 is the bit-exact parity of
tests/test_{mhrs,dcs,ecs,golden,chain,edges}_gpu.py):
  tier 2  conditional E[N_ij | y], E[Z_i | y] and exit-state probabilities against the analytic Hobolth-Jensen values,
          with samples large enough (2e6 paths) to resolve 1e-3;
  tier 3  posterior means and an upper quantile of the engine's chain (through the drop-in LJMA_Gibbs) against the
          reference's own LJMA_Gibbs chain run in the build container and committed as a fixture
          (tests/golden/chains/tier3_reference_chain.npz, made by tests/golden/make_tier3.py)."""
import os

import numpy as np
import pytest

from tests.test_oracle import hobolth_jensen

pytestmark = pytest.mark.gpu

SM = np.array([[-4.1, 1.8, 1.8], [9.5, -11.3, 0.0], [9.5, 0.0, -15.5]])          # SURVEY 8(c), unequal exit rates


def _gpu_means(method, censored, mhit, l=2_000_000, y0=1.5):
    import phasetype_b200 as pb
    from tests import util
    s = -SM.sum(1)
    R = SM.copy(); np.fill_diagonal(R, 0.0)
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    y = np.full(l, y0); cens = np.full(l, 1 if censored else 0, dtype=np.int32)
    eng = pb.Engine(3, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method={"MHRS": 1, "ECS": 2, "DCS": 4}[method], mhit=mhit, seed=31)
    eng.set_theta(theta, next_iter=1)
    N, B, zfix = eng.sweep_stats()
    zbits = eng.zbits
    eng.close()
    return N.reshape(3, 3, order="F") / l, zfix / 2.0 ** zbits / l, s


@pytest.mark.parametrize("method,censored,mhit", [("ECS", False, 1), ("DCS", False, 1), ("ECS", True, 1), ("MHRS", True, 1), ("MHRS", False, 60)])
def test_tier2_conditional_expectations(method, censored, mhit):
    """MHRS on exact data is an independence Metropolis-Hastings chain restarted every sweep: it reaches the conditional
    law only as mhit grows, hence mhit = 60 there; censored MHRS, ECS and DCS are exact samplers."""
    Nm, zm, s = _gpu_means(method, censored, mhit, l=(400_000 if mhit > 1 else 2_000_000))
    Ez, EN, exit_p = hobolth_jensen(SM, s, 1.5, censored)
    off = ~np.eye(3, dtype=bool)
    tol = 3e-3 if method != "MHRS" or censored else 8e-3
    assert np.abs(zm - Ez).max() < tol
    assert np.abs(Nm[off] - EN[off]).max() < 2 * tol
    assert np.abs(np.diag(Nm) - exit_p).max() < tol


@pytest.mark.parametrize("method", [1, 2, 4])
def test_tier3_posterior_against_the_reference_chain(method):
    import phasetype_b200 as pb
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "chains", "tier3_reference_chain.npz"))
    os.environ["PHT_B200_SEED"] = "2718"; os.environ["PHT_B200_QUIET"] = "1"
    it = int(g["it"])
    res = pb.ljma_gibbs(it, 1, method, 3, 2, g["nu"], g["zeta"], g["T"], g["C"], g["y"], g["cens"], [-1.0])
    ref = g["chain_%d" % method]
    for v in range(2):
        xa, xb = res[200:, v], ref[200:, v]
        se = np.sqrt(xa.var() / 100 + xb.var() / 100)        # ~100 effective draws each (conservative)
        assert abs(xa.mean() - xb.mean()) < 4 * se
        assert abs(np.quantile(xa, 0.9) - np.quantile(xb, 0.9)) < 8 * se
