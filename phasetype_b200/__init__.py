"""phasetype_b200 -- B200-native Gibbs engine behind PhaseType's `LJMA_Gibbs`.

Python here is only the host-side mirror of the R wrappers (`phtMCMC`, `phtMCMC2`) and a
ctypes binding; the hot path is CUDA in libpht_b200.so (see include/pht_b200.h).
"""
from .api import phtMCMC, phtMCMC2, ljma_gibbs, METHOD_KEY  # noqa: F401
from ._lib import Engine, EngineError, fp64_fma_rate, lib  # noqa: F401

__all__ = ["phtMCMC", "phtMCMC2", "ljma_gibbs", "Engine", "EngineError", "fp64_fma_rate", "lib", "METHOD_KEY"]
