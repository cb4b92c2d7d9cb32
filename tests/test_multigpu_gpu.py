"""Several GPUs of one box (skipped where fewer than two are visible): the chain must be bit-identical to the single
process CPU restatement whatever the number of devices, (a) through the drop-in LJMA_Gibbs, which fans out over
host threads and the engines' peer-exchange windows (statistics all-reduced by the engine's own peer-memory kernel,
or by NCCL when PHT_B200_NCCL=1 -- both arms are run), and (b) with the global MHRS tail forced
to run from the first tail round on, so that the gather / peer barrier / replicated state machines are exercised on
small data."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from phasetype_b200 import synth

pytestmark = pytest.mark.gpu

CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}


def _ndev():
    import phasetype_b200 as pb
    return pb.lib().pht_device_count()


@pytest.mark.parametrize("method,cid,l,kswitch,cap,mhit", [("MHRS", 3, 60000, "512", "8", 1), ("MHRS", 2, 30000, "512", "4", 3),
                                                           ("MHRS", 3, 300000, "", "", 1), ("ECS", 2, 8000, "", "", 1), ("ECS", 4, 2000, "", "", 1),
                                                           ("DCS", 2, 8000, "", "", 1)])
def test_ljma_gibbs_on_all_devices_equals_oracle(method, cid, l, kswitch, cap, mhit, monkeypatch):
    import phasetype_b200 as pb
    nd = _ndev()
    if nd < 2:
        pytest.skip("needs two or more GPUs")
    gpus = min(nd, 8)
    wl = synth.config(cid, method, l=l)
    monkeypatch.setenv("PHT_B200_SEED", "4242"); monkeypatch.setenv("PHT_B200_SEED_EXACT", "1")
    monkeypatch.setenv("PHT_B200_QUIET", "1"); monkeypatch.setenv("PHT_B200_GPUS", str(gpus))
    if kswitch:
        monkeypatch.setenv("PHT_B200_KSWITCH", kswitch)
    if cap:
        monkeypatch.setenv("PHT_B200_MHRS_CAP", cap)
    it = 5
    got = pb.ljma_gibbs(it, mhit, CODE[method], wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True)
    want, _ = po.gibbs(4242, it, mhit, CODE[method], wl.n, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("method,cid,l", [("MHRS", 3, 60000), ("DCS", 2, 8000)])
def test_nccl_arm_equals_peer_arm(method, cid, l, monkeypatch):
    import phasetype_b200 as pb
    nd = _ndev()
    if nd < 2:
        pytest.skip("needs two or more GPUs")
    wl = synth.config(cid, method, l=l)
    monkeypatch.setenv("PHT_B200_SEED", "99"); monkeypatch.setenv("PHT_B200_SEED_EXACT", "1"); monkeypatch.setenv("PHT_B200_QUIET", "1")
    monkeypatch.setenv("PHT_B200_GPUS", str(min(nd, 8)))
    chains = []
    for arm in ("0", "1"):
        monkeypatch.setenv("PHT_B200_NCCL", arm)
        chains.append(pb.ljma_gibbs(6, 1, CODE[method], wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True))
    assert np.isfinite(chains[0]).all() and (chains[0][1:] > 0).all()
    assert np.array_equal(chains[0], chains[1])


def test_device_count_does_not_change_the_chain(monkeypatch):
    import phasetype_b200 as pb
    nd = _ndev()
    if nd < 2:
        pytest.skip("needs two or more GPUs")
    wl = synth.config(3, "MHRS", l=200000)
    monkeypatch.setenv("PHT_B200_SEED", "77"); monkeypatch.setenv("PHT_B200_SEED_EXACT", "1"); monkeypatch.setenv("PHT_B200_QUIET", "1")
    monkeypatch.setenv("PHT_B200_KSWITCH", "2048")
    chains = []
    for g in sorted({1, 2, min(nd, 8)}):
        monkeypatch.setenv("PHT_B200_GPUS", str(g))
        chains.append(pb.ljma_gibbs(4, 1, 1, wl.n, wl.m, wl.nu, wl.zeta, wl.T, wl.C, wl.y, wl.censored, wl.theta, silent=True))
    for c in chains[1:]:
        assert np.array_equal(c, chains[0])
