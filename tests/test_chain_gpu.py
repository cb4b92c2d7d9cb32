"""Whole-sweep parity on the GPU: assemble -> spectral solve -> paths -> (all-reduce) -> Gamma update, many
sweeps, through the engine API and through the drop-in LJMA_Gibbs.  The CPU restatement (oracle/pht_oracle.c,
pho_gibbs) runs the same Philox contract, the same spectral solver source (pht_eigen.h) and the same
fixed-point sojourn totals, so the chains must agree to the last bit."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from phasetype_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu

CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}

# the two fixtures of the reference's own tests (tests/phtMCMC.R, tests/phtMCMC2.R) as raw LJMA_Gibbs arguments
Y20 = np.array([1.45353415045187, 1.85349532001349, 2.01084961814576, 0.505725921290172, 1.56252630012213,
                3.41158665930278, 1.52674487509487, 4.3428662377235, 8.03208018151311, 2.41746547476986,
                0.38828086509283, 2.61513815012196, 3.39148865480856, 1.82705817807965, 1.42090953713845,
                0.851438991331866, 0.0178808867191894, 0.632198596390046, 0.959910259815998, 1.83344199966323])
FIX1 = dict(it=6, mhit=1, method=1, n=3, T=[0, 3, 5, 0, 1, 0, 6, 0, 2, 4, 0, 0, 7, 8, 9, 0], C=np.ones(16),
            nu=[24, 24, 180, 1, 180, 1, 1, 24, 24], zeta=[16.0] * 9)
FIX2 = dict(it=20, mhit=1, method=2, n=3, T=[0, 2, 2, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0], C=np.ones(16),
            nu=[24, 180], zeta=[16.0, 16.0])


@pytest.mark.parametrize("method,cid,l,sweeps", [("MHRS", 3, 5000, 6), ("MHRS", 2, 3000, 5), ("DCS", 3, 3000, 5),
                                                 ("ECS", 3, 3000, 5), ("ECS", 4, 2000, 4), ("DCS", 1, 100, 30)])
def test_engine_chain_equals_oracle(method, cid, l, sweeps):
    import phasetype_b200 as pb
    wl = synth.config(cid, method, l=l)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=CODE[method], mhit=1, seed=2024,
                    mhrs_cap=16)
    eng.set_theta(wl.theta, next_iter=1)
    got = eng.run(sweeps)
    eng.close()
    want, _ = po.gibbs(2024, sweeps + 1, 1, CODE[method], wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    assert np.array_equal(got, want[1:])


@pytest.mark.parametrize("fx,graph", [(FIX1, "1"), (FIX2, "1"), (FIX2, "0")])
def test_ljma_gibbs_reference_fixtures(fx, graph):
    """The drop-in routine on the argument vectors R would build for the reference's two test scripts."""
    import phasetype_b200 as pb
    os.environ["PHT_B200_SEED"] = "12345"; os.environ["PHT_B200_QUIET"] = "1"; os.environ["PHT_B200_GRAPH"] = graph
    m = len(fx["nu"])
    cens = np.zeros(20, dtype=np.int32)
    res = pb.ljma_gibbs(fx["it"], fx["mhit"], fx["method"], fx["n"], m, fx["nu"], fx["zeta"], fx["T"], fx["C"], Y20, cens, [-1.0])
    want, _ = po.gibbs(12345, fx["it"], fx["mhit"], fx["method"], fx["n"], fx["nu"], fx["zeta"], fx["T"], fx["C"], Y20, cens, [-1.0])
    assert res.shape == (fx["it"], m)
    assert np.array_equal(res, want)
    assert (res > 0).all()


def test_api_mirror_matches_raw_call():
    """phtMCMC2 (Python mirror of R/phtMCMC2.R) builds the same vectors as fixture 2 and resumes like R does."""
    import phasetype_b200 as pb
    os.environ["PHT_B200_SEED"] = "99"; os.environ["PHT_B200_QUIET"] = "1"; os.environ["PHT_B200_GRAPH"] = "1"
    TT = np.array([["0", "F", "F", "0"], ["R", "0", "0", "F"], ["R", "0", "0", "F"], ["0", "0", "0", "0"]], dtype=object)
    out = pb.phtMCMC2(Y20, TT, [1, 0, 0], {"R": 180, "F": 24}, {"R": 16, "F": 16}, 20, silent=True)
    raw = pb.ljma_gibbs(20, 1, 2, 3, 2, FIX2["nu"], FIX2["zeta"], FIX2["T"], FIX2["C"], Y20, np.zeros(20, dtype=np.int32), [-1.0])
    assert out["vars"] == ["F", "R"] and np.array_equal(out["samples"], raw)
    more = pb.phtMCMC2(Y20, TT, [1, 0, 0], {"R": 180, "F": 24}, {"R": 16, "F": 16}, 5, resume=out, silent=True)
    assert more["samples"].shape == (25, 2) and np.array_equal(more["samples"][:20], raw)
    dense = pb.phtMCMC(Y20, 3, [1, 0, 0], [24, 24, 1, 180, 1, 24, 180, 1, 24], [16, 16, 16], 6, mhit=1, silent=True)
    assert dense["samples"].shape == (6, 9) and dense["vars"] == ["S12", "S13", "S21", "S23", "S31", "S32", "s1", "s2", "s3"]
    want, _ = po.gibbs(99, 6, 1, 1, 3, FIX1["nu"], FIX1["zeta"], FIX1["T"], FIX1["C"], Y20, np.zeros(20, dtype=np.int32), [-1.0])
    assert np.array_equal(dense["samples"], want)
