"""The CUDA kernels against the committed golden vectors (dumps of the unmodified reference C)."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
CODE = {"MHRS": 1, "ECS": 2, "DCS": 4}


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_kernels_reproduce_golden_vectors(path):
    import phasetype_b200 as pb
    g = np.load(path)
    n = int(g["n"]); S = g["S"].reshape(n, n, order="F"); s = g["s"]
    # parameters = the rates themselves, every non-zero cell its own variable; the diagonal the engine assembles
    # (ascending row sum) must reproduce the golden S exactly, else the vector is not usable for a bit test
    T = np.zeros((n + 1, n + 1), dtype=np.int32); theta = []
    for j in range(n + 1):
        for i in range(n):
            v = s[i] if j == n else S[i, j]
            if i != j and v != 0.0:
                theta.append(v); T[i, j] = len(theta)
    theta = np.array(theta); m = theta.shape[0]
    eng = pb.Engine(n, T.ravel(order="F"), np.ones((n + 1) ** 2), np.full(m, 2.0), np.full(m, 2.0), g["y"], g["cens"],
                    method=CODE[str(g["method"])], mhit=int(g["mhit"]), seed=int(g["seed"]), mhrs_cap=8)
    if str(g["method"]) != "MHRS":
        eng.set_spectral(g["evals"], g["Q"], g["Qinv"])
    eng.set_theta(theta, next_iter=int(g["it"]))
    B, N, z = eng.paths()
    mdl = eng.model()
    eng.close()
    assert np.array_equal(mdl["S"], g["S"]), "assembled generator differs from the golden one"
    assert np.array_equal(B, g["B"]) and np.array_equal(N, g["N"]) and np.array_equal(z, g["z"])
