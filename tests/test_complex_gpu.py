"""Spectra with complex pairs.  The reference takes the real parts of LAPACK's output and carries on
(src/utility.c:116-121), i.e. its spectral samplers are silently wrong on a general dense generator; the engine's solver
returns the real block form S = Q B Q^-1 (pht_eigen.h) and ECS evaluates exp(xS) through it.  There is nothing in the
reference to be bit-identical to here, so the checks are (a) the decomposition itself against numpy, on the device,
and (b) tier 2: conditional E[N_ij | y], E[Z_i | y] and exit-state probabilities of the CUDA sampler against the
analytic Hobolth-Jensen values, for exact and for right-censored observations (ECS) and for exact ones through DCS
and the MH variants, whose DCS-family kernels run their CPLX instance (block formulas in complex arithmetic) on such a sweep."""
import numpy as np
import pytest

from tests import util
from tests.test_oracle import hobolth_jensen

pytestmark = pytest.mark.gpu

# an unsymmetric 4-phase generator with one complex pair and unequal exit rates (cyclic flow 0 -> 1 -> 2 -> 3 -> 0)
R4 = np.array([[0.0, 2.6, 0.1, 0.2], [0.15, 0.0, 2.9, 0.1], [0.2, 0.1, 0.0, 2.4], [2.2, 0.2, 0.1, 0.0]])
S4 = np.array([0.5, 0.2, 0.9, 0.3])


def _engine(method, y, cens, R=R4, s=S4, seed=17):
    import phasetype_b200 as pb
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    eng = pb.Engine(s.shape[0], T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=method, mhit=1, seed=seed)
    eng.set_theta(theta, next_iter=1)
    return eng, T, C, theta


def test_device_block_decomposition():
    import phasetype_b200 as pb
    rng = np.random.default_rng(5)
    for n in (4, 8, 16, 32):
        R, s = (R4, S4) if n == 4 else util.dense_rates(n, rng)
        eng, T, C, theta = _engine(2, np.array([1.0]), np.zeros(1, dtype=np.int32), R, s)
        eng.sweep_stats()
        mdl = eng.model()
        eng.close()
        Sm = mdl["S"].reshape(n, n, order="F"); Q = mdl["Q"].reshape(n, n, order="F"); Qi = mdl["Qinv"].reshape(n, n, order="F")
        lam = np.linalg.eigvals(Sm)
        assert np.abs(lam.imag).max() > 1e-3                      # the point of the test
        assert np.allclose(np.sort(mdl["evals"]), np.sort(lam.real), atol=1e-10 * np.abs(lam).max())
        assert np.abs(Q @ Qi - np.eye(n)).max() < 1e-9
        B = Qi @ Sm @ Q                                            # must be block diagonal: 1x1 and [[a, b], [-b, a]] blocks
        mask = np.eye(n, dtype=bool)
        k = 0
        while k < n:
            if k + 1 < n and abs(B[k, k + 1]) > 1e-9:
                assert abs(B[k, k] - B[k + 1, k + 1]) < 1e-9 and abs(B[k, k + 1] + B[k + 1, k]) < 1e-9 and B[k, k + 1] > 0
                mask[k, k + 1] = mask[k + 1, k] = True; k += 2
            else:
                k += 1
        assert np.abs(B[~mask]).max() < 1e-9 * np.abs(lam).max() * n


@pytest.mark.parametrize("censored", [False, True])
def test_ecs_tier2_on_a_complex_spectrum(censored):
    l, y0 = 1_000_000, 1.3
    y = np.full(l, y0); cens = np.full(l, 1 if censored else 0, dtype=np.int32)
    eng, T, C, theta = _engine(2, y, cens)
    N, B, zfix = eng.sweep_stats()
    zbits = eng.zbits
    cnt = eng.counters()
    eng.close()
    n = 4
    Nm = N.reshape(n, n, order="F") / l; zm = zfix / 2.0 ** zbits / l
    S, s = util.assemble(T, C, theta, n)
    Sm = S.reshape(n, n, order="F")
    assert np.abs(np.linalg.eigvals(Sm).imag).max() > 0.5
    Ez, EN, exit_p = hobolth_jensen(Sm, s, y0, censored)
    off = ~np.eye(n, dtype=bool)
    assert np.abs(zm - Ez).max() < 4e-3
    assert np.abs(Nm[off] - EN[off]).max() < 8e-3
    assert np.abs(np.diag(Nm) - exit_p).max() < 4e-3
    assert cnt["nonfinite"] == 0


def test_ecs_chain_runs_on_the_general_dense_workload():
    """BASELINE config 3 as written (8-phase GENERAL dense generator, 20 % censored) under ECS: the posterior draws
    wander through generators with complex pairs; the chain must stay finite and positive and near the truth."""
    import phasetype_b200 as pb
    from phasetype_b200 import synth
    wl = synth.config(3, "MHRS", l=20000)          # the unsymmetrised variant of the shape
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=2, seed=3)
    eng.set_theta(wl.theta, next_iter=1)
    out = eng.run(30)
    eng.close()
    assert np.isfinite(out).all() and (out > 0).all()
    assert np.abs(np.log(out[10:].mean(0) / wl.theta)).max() < 1.0


@pytest.mark.parametrize("method,mhit", [(4, 1), (8, 30), (16, 30)])
def test_dcs_family_tier2_on_a_complex_spectrum(method, mhit):
    """DCS (method 4) and the two MH variants (8: Hobolth chains, 16: Aslett chains; enough proposals to reach the
    conditional law given absorption AT y) on a generator with a complex pair: analytic conditional expectations."""
    import phasetype_b200 as pb
    l, y0 = 400_000, 1.3
    y = np.full(l, y0); cens = np.zeros(l, dtype=np.int32)
    T, C, theta = util.general_model(R4, S4)
    m = theta.shape[0]
    eng = pb.Engine(4, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=method, mhit=mhit, seed=23)
    eng.set_theta(theta, next_iter=1)
    N, B, zfix = eng.sweep_stats()
    zbits = eng.zbits
    cnt = eng.counters()
    eng.close()
    n = 4
    Nm = N.reshape(n, n, order="F") / l; zm = zfix / 2.0 ** zbits / l
    S, s = util.assemble(T, C, theta, n)
    Sm = S.reshape(n, n, order="F")
    assert np.abs(np.linalg.eigvals(Sm).imag).max() > 0.5
    Ez, EN, exit_p = hobolth_jensen(Sm, s, y0, False)
    off = ~np.eye(n, dtype=bool)
    assert B.sum() == l and cnt["nonfinite"] == 0
    assert np.abs(zm - Ez).max() < 6e-3
    assert np.abs(Nm[off] - EN[off]).max() < 1.2e-2
    assert np.abs(np.diag(Nm) - exit_p).max() < 6e-3


def test_dcs_chain_runs_on_the_general_dense_workload():
    """BASELINE config 3 as written under DCS: the chain wanders through generators with and without complex pairs."""
    import phasetype_b200 as pb
    from phasetype_b200 import synth
    wl = synth.config(3, "MHRS", l=20000)
    cens = np.zeros_like(wl.censored)              # DCS ignores the flag; give it exact observations
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, cens, method=4, seed=3)
    eng.set_theta(wl.theta, next_iter=1)
    out = eng.run(20)
    eng.close()
    assert np.isfinite(out).all() and (out > 0).all()
