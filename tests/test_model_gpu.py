"""The once-per-sweep model work on the device against independent checkers: the engine's spectral solver
(k_spectral_solve, shared source pht_eigen.h) must equal the host build of the same source bit for bit AND satisfy
the invariants LAPACK's decomposition satisfies (reference src/utility.c:87-129 uses dgeevx + dgetrf/dgetri), so the
whole-chain parity of the spectral samplers is not solver-versus-same-solver only."""
import numpy as np
import pytest

from oracle import pyoracle as po
from phasetype_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cid,method,code", [(1, "DCS", 4), (2, "ECS", 2), (3, "DCS", 4), (3, "ECS", 2), (4, "ECS", 2), (5, "DCS", 4)])
def test_device_spectral_solver(cid, method, code):
    import phasetype_b200 as pb
    wl = synth.config(cid, method, l=64)
    n = wl.n
    eng = pb.Engine(n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=code, seed=3)
    eng.set_theta(wl.theta, next_iter=1)
    eng.sweep_stats()                          # runs k_assemble + k_spectral_solve
    mdl = eng.model()
    eng.close()
    S, s = util.assemble(wl.T, wl.C, wl.theta, n)
    assert np.array_equal(mdl["S"], S) and np.array_equal(mdl["s"], s)
    P, Pfull = po.embedded(S, s)
    assert np.array_equal(mdl["P"], P, equal_nan=True) and np.array_equal(mdl["Pfull"], Pfull, equal_nan=True)
    # (1) bit-identical to the host build of the same solver source
    ev, Q, Qi = po.eigen("native", S, n)
    assert np.array_equal(mdl["evals"], ev) and np.array_equal(mdl["Q"], Q) and np.array_equal(mdl["Qinv"], Qi)
    # (2) the invariants of the decomposition, against numpy's LAPACK
    Sm = S.reshape(n, n, order="F"); Qm = mdl["Q"].reshape(n, n, order="F"); Qim = mdl["Qinv"].reshape(n, n, order="F")
    lam = np.linalg.eigvals(Sm)
    assert np.abs(lam.imag).max() < 1e-9
    scale = np.abs(lam).max()
    assert np.allclose(np.sort(mdl["evals"]), np.sort(lam.real), rtol=0, atol=1e-11 * scale)
    assert np.abs(Qm @ Qim - np.eye(n)).max() < 1e-9
    assert np.abs(Qm @ np.diag(mdl["evals"]) @ Qim - Sm).max() < 1e-10 * scale * n
    # exp(S t) through the device's spectral data equals scipy's expm
    from scipy.linalg import expm
    for t in (0.1, 1.0, 5.0):
        assert np.allclose(Qm @ np.diag(np.exp(mdl["evals"] * t)) @ Qim, expm(Sm * t), rtol=0, atol=1e-10)
