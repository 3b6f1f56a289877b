"""psmc_b200 -- B200-native E-step for the PSMC HMM (drop-in for lh3/psmc's em.c:33-55 / aux.c:150-201).

The product is libpsmc_b200.so (CUDA kernels behind the C ABI of include/psmc_b200.h) and the `psmc`
host driver under host/.  This package is the thin Python mirror of that C ABI used by the tests,
bench.py and the multi-GPU driver; it holds no numerics of its own and has NO CPU fallback.
"""
from ._lib import load_library, LibraryNotBuilt, Psmc200Error  # noqa: F401
from .estep import EStep, Model, factorize  # noqa: F401

__version__ = "0.1.0"
