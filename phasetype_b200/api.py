"""Python mirror of the reference's R wrappers, so the drop-in routine can be exercised
without R (not installed in this image).

`phtMCMC2` / `phtMCMC` reproduce the argument checks, variable naming, TN encoding and
result shaping of reference R/phtMCMC2.R:1-86 and R/phtMCMC.R:1-97 and then make the same
native call R makes -- `.C(LJMA_Gibbs, it, mhit, method, n, m, nu, zeta, T, C, y, l, censored,
start, silent, res)` (R/phtMCMC2.R:73) -- into libpht_b200.so.
"""
import numpy as np

from . import _lib

METHOD_KEY = {"MHRS": 1, "ECS": 2, "DCS": 4}   # R/phtMCMC2.R:66


def ljma_gibbs(it, mhit, method, n, m, nu, zeta, T, C, y, censored, start, silent=True):
    """The raw 15-argument native call; returns res as an (it, m) array (R: matrix(res$res, nrow=n))."""
    L = _lib.lib()
    i32 = lambda v: np.array([v], dtype=np.int32)
    nu = np.ascontiguousarray(nu, dtype=np.float64); zeta = np.ascontiguousarray(zeta, dtype=np.float64)
    T = np.ascontiguousarray(np.asarray(T, dtype=np.int32).ravel()); C = np.ascontiguousarray(np.asarray(C, dtype=np.float64).ravel())
    y = np.ascontiguousarray(y, dtype=np.float64); censored = np.ascontiguousarray(censored, dtype=np.int32)
    start = np.ascontiguousarray(np.atleast_1d(start), dtype=np.float64)
    if start.shape[0] < m:
        start = np.concatenate([start, np.zeros(m - start.shape[0])])
    res = np.zeros(it * m, dtype=np.float64)
    L.LJMA_Gibbs(i32(it), i32(mhit), i32(method), i32(n), i32(m), nu, zeta, T, C, y, i32(y.shape[0]), censored,
                 start, i32(1 if silent else 0), res)
    return res.reshape(m, it).T.copy()     # column-major it x m


def _encode(TT, nu, zeta):
    """Variable names -> sorted order -> integer matrix TN (R/phtMCMC2.R:27-63). C-locale sort."""
    TT = np.asarray(TT, dtype=object)
    dimT = TT.shape[0]
    names = sorted({str(v) for v in TT.ravel()} - {"0"})
    if sorted(nu.keys()) != names:
        raise ValueError("variables specified in matrix don't match those in prior nu")
    if sorted(zeta.keys()) != names:
        raise ValueError("variables specified in matrix don't match those in prior zeta")
    index = {nm: k + 1 for k, nm in enumerate(names)}
    TN = np.zeros((dimT, dimT), dtype=np.int32)
    for i in range(dimT):
        for j in range(dimT):
            TN[i, j] = index.get(str(TT[i, j]), 0)
    return names, TN


def phtMCMC2(x, TT, beta, nu, zeta, n, censored=None, C=None, method="ECS", mhit=1, resume=None, silent=False):
    """Mirror of R/phtMCMC2.R.  TT: (k x k) array of variable names with "0"/0 for structural zeros;
    nu, zeta: dicts name -> value; resume: dict with "samples" (array) and "vars" from a previous run."""
    x = np.asarray(x, dtype=np.float64)
    TT = np.asarray(TT, dtype=object)
    if not (isinstance(n, (int, np.integer)) and n >= 1):
        raise ValueError("%r is an invalid number of MCMC iterations." % (n,))
    if not (isinstance(mhit, (int, np.integer)) and mhit >= 0):
        raise ValueError("%r is an invalid number of Metropolis-Hastings iterations." % (mhit,))
    if TT.ndim != 2 or TT.shape[0] != TT.shape[1]:
        raise ValueError("matrix of variables must be square")
    dimT = TT.shape[0]
    if any(str(TT[i, i]) != "0" for i in range(dimT)):
        raise ValueError("diagonal of matrix of variables must be zeros")
    if any(str(v) != "0" for v in TT[dimT - 1, :]):
        raise ValueError("last row of matrix of variables must represent absorbing state (and so be all zeros)")
    beta = np.asarray(beta, dtype=np.float64)
    if beta.shape[0] != dimT - 1:
        raise ValueError("beta should be a vector of length %d for the generator specified." % (dimT - 1))
    if (beta < 0).any():
        raise ValueError("beta is not a valid parameter of a Dirichlet distribution.")
    if C is None:
        C = np.ones((dimT, dimT))
    C = np.asarray(C, dtype=np.float64)
    if C.shape != (dimT, dimT):
        raise ValueError("dimension of C must match dimension of TT")
    if censored is None:
        censored = np.zeros(x.shape[0], dtype=bool)
    nu = dict(nu); zeta = dict(zeta)
    names, TN = _encode(TT, nu, zeta)
    start = np.array([-1.0])
    prev = None
    if resume is not None:
        if list(resume["vars"]) != names:
            raise ValueError("the variable names in resume do not match the variable names in generator")
        prev = np.asarray(resume["samples"], dtype=np.float64)
        start = prev[-1, :].copy()
        n = n + 1
    methods = [method] if isinstance(method, str) else list(method)
    unknown = [mm for mm in methods if mm not in METHOD_KEY]
    if unknown:
        raise ValueError("Error: unknown sampling methods (%s)" % ", ".join(unknown))
    method_num = sum(METHOD_KEY[mm] for mm in set(methods))
    res = ljma_gibbs(int(n), int(mhit), method_num, dimT - 1, len(names),
                     [nu[k] for k in names], [zeta[k] for k in names],
                     TN.ravel(order="F"), C.ravel(order="F"), x, np.asarray(censored).astype(np.int32), start, silent)
    samples = res if prev is None else np.vstack([prev[:-1, :], res])
    return {"samples": samples, "data": x, "vars": names, "TT": TT, "beta": beta, "nu": nu, "zeta": zeta,
            "iterations": n, "censored": np.asarray(censored), "method": method, "MHit": mhit}


def phtMCMC(x, states, beta, nu, zeta, n, mhit=1, resume=None, silent=False):
    """Mirror of R/phtMCMC.R: dense `states`-phase generator, nu row-wise over (S_i., s_i), one zeta per row,
    method fixed to MHRS and no censoring (R/phtMCMC.R:3-4).  Names use a separator-free paste as in R
    (R/phtMCMC.R:17), so states >= 11 collide exactly as upstream does."""
    nu = np.asarray(nu, dtype=np.float64); zeta = np.asarray(zeta, dtype=np.float64)
    if nu.shape[0] != states * states:
        raise ValueError("nu must specify one prior Gamma shape parameter per element of the Phase-type generator matrix")
    if zeta.shape[0] != states:
        raise ValueError("zeta must specify one prior Gamma reciprocal scale parameter per non-absorbing row")
    TT = np.empty((states + 1, states + 1), dtype=object)
    TT[:, :] = "0"
    for i in range(states):
        for j in range(states):
            if i != j:
                TT[i, j] = "S%d%d" % (i + 1, j + 1)
        TT[i, states] = "s%d" % (i + 1)
    rowwise = [str(TT[i, j]) for i in range(states + 1) for j in range(states + 1) if str(TT[i, j]) != "0"]
    if len(set(rowwise)) != len(rowwise):
        raise ValueError("variables specified in matrix don't match those in prior nu")
    nu_d = dict(zip(rowwise, nu))
    # R: `zeta <- as.list(rep(zeta, each=states+1)); names(zeta) <- <states^2 row-wise names>` (R/phtMCMC.R:29-30):
    # the k-th name is paired with element k of the (states+1)-fold repetition -- reproduced literally
    zeta_d = {nm: zeta[k // (states + 1)] for k, nm in enumerate(rowwise)}
    out = phtMCMC2(x, TT, beta, nu_d, zeta_d, n, method="MHRS", mhit=mhit, resume=resume, silent=silent)
    return out
