"""Edge cases of the path on the GPU, each checked against the oracle: a single phase, single and zero observations,
every observation censored, an empty shard, and the L2-flush aid (complex spectra: tests/test_complex_gpu.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from phasetype_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu

CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}


def _paths(method, R, s, y, cens, seed=11, it=3, mhit=1, cap=0):
    import phasetype_b200 as pb
    n = s.shape[0]
    T, C, theta = util.general_model(R, s)
    m = theta.shape[0]
    eng = pb.Engine(n, T, C, np.full(m, 2.0), np.full(m, 2.0), y, cens, method=CODE[method], mhit=mhit, seed=seed, mhrs_cap=cap)
    S, sv = util.assemble(T, C, theta, n)
    spec = None
    if method != "MHRS":
        spec = po.eigen("oracle", S, n)
        eng.set_spectral(*spec)
    eng.set_theta(theta, next_iter=it)
    B, N, z = eng.paths()
    eng.close()
    if method == "MHRS":
        Bo, No, zo, _ = po.mhrs_paths("oracle", seed, it, y, cens, S, sv, mhit=mhit)
    else:
        Bo, No, zo, _ = po.spectral_paths("oracle", method, seed, it, y, cens, S, sv, spectral=spec)
    return (B, N, z), (Bo, No, zo)


@pytest.mark.parametrize("method", ["MHRS", "ECS", "DCS"])
def test_single_phase_is_the_exponential_distribution(method):
    """n = 1: the only path is 'stay in state 1 until y' (exact) -- N = [1], z = [y]."""
    R = np.zeros((1, 1)); s = np.array([0.7])
    rng = np.random.default_rng(1)
    y = rng.exponential(1.0 / 0.7, 300) + 1e-3
    cens = np.zeros(300, dtype=np.int32)
    got, want = _paths(method, R, s, y, cens)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    assert (got[1] == 1).all() and np.array_equal(got[2][:, 0], y)


@pytest.mark.parametrize("method", ["MHRS", "ECS", "DCS"])
def test_one_observation(method):
    rng = np.random.default_rng(2)
    R, s = util.dense_rates(4, rng, symmetric=True)
    got, want = _paths(method, R, s, np.array([0.83]), np.zeros(1, dtype=np.int32))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("method", ["MHRS", "ECS", "DCS"])
def test_every_observation_censored(method):
    """MHRS and ECS draw paths absorbed after y; DCS ignores the flag (src/Simulate_AbsCTMC_eq_AslettHobolth_DCS.c:132-133)."""
    rng = np.random.default_rng(3)
    R, s = util.dense_rates(5, rng, symmetric=True)
    y = rng.exponential(0.8, 800) + 0.01
    cens = np.ones(800, dtype=np.int32)
    got, want = _paths(method, R, s, y, cens, cap=8)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    if method != "DCS":
        assert (got[2].sum(1) >= y - 1e-12).all()          # the latent chain lives at least until the censoring time
    else:
        assert np.allclose(got[2].sum(1), y, rtol=1e-12)


@pytest.mark.parametrize("method", ["MHRS", "DCS", "ECS"])
def test_zero_observations_and_empty_shard(method):
    """No data: the sweep statistics are zero and the update is a prior draw; a rank whose shard is empty still runs."""
    import phasetype_b200 as pb
    # prior draws are not symmetric, so the spectral samplers get the Coxian shape (triangular S: real spectrum always)
    wl = synth.config(3 if method == "MHRS" else 2, method, l=64)
    none = np.zeros(0)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, none, none.astype(np.int32), method=CODE[method], seed=5, sum_y_global=1.0)
    eng.set_theta(wl.theta, next_iter=1)
    N, B, z = eng.sweep_stats()
    assert not N.any() and not B.any() and not z.any()
    got = eng.run(3)
    eng.close()
    want, _ = po.gibbs(5, 4, 1, CODE[method], wl.n, wl.nu, wl.zeta, wl.T, wl.C, none, none.astype(np.int32), wl.theta)
    assert np.array_equal(got, want[1:])
    # world = 3 over 2 observations: rank 2 owns nothing
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y[2:2:3], wl.censored[2:2:3], method=CODE[method], seed=5, rank=2, world=3,
                    sum_y_global=float(wl.y[:2].sum()))
    eng.set_theta(wl.theta, next_iter=1)
    N, B, z = eng.sweep_stats()
    eng.close()
    assert not N.any() and not B.any() and not z.any()


def test_l2_flush_does_not_change_the_chain():
    import phasetype_b200 as pb
    wl = synth.config(2, "MHRS", l=4000)
    out = []
    for flush in (0, 8 << 20):
        eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, seed=77)
        eng.set_l2_flush(flush)
        eng.set_theta(wl.theta, next_iter=1)
        out.append(eng.run(4))
        eng.close()
    assert np.array_equal(out[0], out[1])


def test_full_tail_lists_only_cost_speed(monkeypatch):
    """With room for 8 handed-over observations only, the lanes that find the lists full keep their observation and go
    on by themselves: same paths, bit for bit."""
    monkeypatch.setenv("PHT_B200_TAIL_SLOTS", "8")
    rng = np.random.default_rng(5)
    R, s = util.dense_rates(6, rng)
    y = rng.exponential(1.5, 4000) + 0.01
    cens = (rng.uniform(size=4000) < 0.2).astype(np.int32)
    got, want = _paths("MHRS", R, s, y, cens, mhit=2, cap=4)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_round_trace_counts_the_speculation():
    """Per tail round the trace holds the attempts run next to the attempts the sequential sampler needs of that round's
    offer: the parallel search can only run MORE (whatever completes beyond a survivor before it is known)."""
    import phasetype_b200 as pb
    wl = synth.config(2, "MHRS", l=20000)
    eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, mhit=1, seed=3, mhrs_cap=4, use_graph=False)
    eng.set_theta(wl.theta, next_iter=1)
    eng.run(2)
    tr = eng.round_trace(); cnt = eng.counters()
    eng.close()
    assert tr.shape == (48, 8)
    used = tr[:, 4] > 0
    assert used.sum() >= 2 and cnt["tail_rounds"] >= used.sum()
    assert (tr[used, 6] > 0).all() and (tr[used, 5] >= tr[used, 6]).all()       # run >= needed > 0
    assert (tr[used, 7] >= tr[used, 5]).all()                                   # an attempt is at least one jump-step
    assert not tr[~used].any()


def test_cached_device_memory_can_be_released_between_calls():
    """The large buffers of a call stay cached in the device's memory pool; handing them back must not disturb the next call."""
    import phasetype_b200 as pb
    from phasetype_b200 import _lib
    wl = synth.config(2, "MHRS", l=3000)
    runs = []
    for rep in range(3):
        eng = pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y, wl.censored, method=1, mhit=1, seed=11)
        eng.set_theta(wl.theta, next_iter=1)
        runs.append(eng.run(3))
        eng.close()
        if rep == 0:
            assert _lib.lib().pht_release_device_memory() == 0
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])


def test_ranks_of_a_run_must_share_their_configuration():
    """Two engines that claim to be ranks of one run but disagree on the fixed-point scale cannot be attached to one another."""
    import phasetype_b200 as pb
    wl = synth.config(2, "MHRS", l=2000)
    engs = [pb.Engine(wl.n, wl.T, wl.C, wl.nu, wl.zeta, wl.y[r::2], wl.censored[r::2], method=1, mhit=1, seed=5, rank=r, world=2, zbits=30 + r)
            for r in range(2)]
    handles = b"".join(e.peer_handle() for e in engs)
    with pytest.raises(pb.EngineError, match="different configuration"):
        engs[0].peer_attach(handles)
    for e in engs:
        e.close()
