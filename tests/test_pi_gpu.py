"""Start distribution.  The reference hard-wires pi = e1 with a FIX ME where its Dirichlet update belongs
(src/PHT_MCMC_Aslett.c:190-193), but its samplers take pi as an argument.  So: (1) tier 1 for paths started from a
general pi -- CUDA kernels against the restatement AND the unmodified reference C, bit for bit; (2) the conjugate draw
pi | paths ~ Dirichlet(beta + B) itself against a restatement from the same Philox gamma streams, bit for bit;
(3) the posterior of pi concentrates on the truth on data simulated from a known start distribution."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import util

pytestmark = pytest.mark.gpu

CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}


@pytest.mark.parametrize("method,n,fc", [("MHRS", 4, 0.2), ("MHRS", 8, 0.0), ("ECS", 5, 0.3), ("DCS", 6, 0.0)])
def test_paths_from_a_general_start_distribution(method, n, fc):
    import phasetype_b200 as pb
    rng = np.random.default_rng(40 + n)
    R, s = util.dense_rates(n, rng, symmetric=(method != "MHRS"))
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    pi = rng.dirichlet(np.ones(n))
    l = 2500
    y = rng.exponential(1.0, l) + 0.01; cens = (rng.uniform(size=l) < fc).astype(np.int32)
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=CODE[method], mhit=1, seed=99, mhrs_cap=16)
    S, sv = util.assemble(T, C, theta, n)
    spec = None
    if method != "MHRS":
        spec = po.eigen("oracle", S, n); eng.set_spectral(*spec)
    eng.set_pi(pi)
    eng.set_theta(theta, next_iter=2)
    B, N, z = eng.paths()
    assert np.array_equal(eng.get_pi(), pi)
    eng.close()
    po.set_pi(pi)
    try:
        impls = ["oracle"] + (["ref"] if po.have_ref() else [])
        for impl in impls:
            if method == "MHRS":
                Bo, No, zo, _ = po.mhrs_paths(impl, 99, 2, y, cens, S, sv, mhit=1)
            else:
                Bo, No, zo, _ = po.spectral_paths(impl, method, 99, 2, y, cens, S, sv, spectral=spec)
            assert np.array_equal(B, Bo) and np.array_equal(N, No) and np.array_equal(z, zo), impl
    finally:
        po.set_pi(None)
    assert len(np.unique(B)) > 1          # paths really start in several states


def test_dirichlet_draw_equals_its_restatement():
    import phasetype_b200 as pb
    rng = np.random.default_rng(3)
    n = 4
    R, s = util.dense_rates(n, rng)
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    y = rng.exponential(1.0, 5000) + 0.01; cens = np.zeros(5000, dtype=np.int32)
    beta = np.array([1.0, 0.5, 2.0, 1.5]); pi0 = np.array([0.4, 0.3, 0.2, 0.1])
    seed = 12345
    eng = pb.Engine(n, T, C, np.full(m, 2.0), 2.0 / theta, y, cens, method=1, mhit=1, seed=seed)
    eng.set_pi(pi0, beta)
    eng.set_theta(theta, next_iter=1)
    Nst, Bst, zst = eng.sweep_stats()                      # the statistics sweep 1 will see (same key, same parameters)
    out = eng.run(2)
    rows = eng.pi_rows(2)
    eng.close()
    g = np.array([po.rgamma_at(seed, 1, m + i, beta[i] + float(Bst[i]), 1.0) for i in range(n)])
    tot = 0.0
    for v in g:
        tot += v
    assert np.array_equal(rows[0], g / tot)
    assert Bst.sum() == 5000 and (Bst > 0).all()
    assert np.allclose(rows.sum(1), 1.0) and (rows > 0).all() and np.isfinite(out).all()


def test_posterior_of_pi_finds_the_truth():
    import phasetype_b200 as pb
    rng = np.random.default_rng(11)
    n = 3
    # three nearly closed phases with very different exit rates: the absorption time says a lot about the start state, so the
    # data-augmentation chain mixes in a few sweeps (with strongly communicating phases it needs thousands)
    R = np.array([[0.0, 0.02, 0.01], [0.01, 0.0, 0.02], [0.01, 0.01, 0.0]]); s = np.array([0.2, 2.0, 20.0])
    pi_true = np.array([0.6, 0.1, 0.3])
    # data from the model started from pi_true
    rate = R.sum(1) + s; P = np.concatenate([R, s[:, None]], axis=1) / rate[:, None]
    l = 6000
    y = np.zeros(l)
    for k in range(l):
        j = rng.choice(n, p=pi_true); t = 0.0
        while j < n:
            t += rng.exponential(1.0 / rate[j]); j = rng.choice(n + 1, p=P[j])
        y[k] = t
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    # rates pinned by a tight prior at the truth, so that the chain explores pi only
    # MHRS with many MH proposals per sweep: on exact observations it reaches the conditional law as mhit grows
    eng = pb.Engine(n, T, C, np.full(m, 4000.0), 4000.0 / theta, y, np.zeros(l, dtype=np.int32), method=1, mhit=100, seed=5)
    eng.set_pi(np.full(n, 1.0 / n), np.ones(n))
    eng.set_theta(theta, next_iter=1)
    eng.run(400)
    rows = eng.pi_rows(400)
    eng.close()
    post = rows[100:].mean(0)
    # (the residual bias of the restarted independence sampler at mhit = 100 is ~0.02; at mhit = 1 the chain collapses
    # onto the slowest phase -- SURVEY H6 applies to B as it does to N and z)
    assert np.abs(post - pi_true).max() < 0.06


def test_pi_update_is_refused_under_ecs_and_dcs():
    import phasetype_b200 as pb
    from phasetype_b200 import synth
    wl = synth.config(2, "ECS", l=64)
    for code in (2, 4):
        eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=code, seed=1)
        with pytest.raises(pb.EngineError, match="needs method MHRS"):
            eng.set_pi(None, np.ones(wl.n))
        eng.set_pi(np.full(wl.n, 1.0 / wl.n))          # a fixed general pi is fine (reference draw order)
        eng.close()
