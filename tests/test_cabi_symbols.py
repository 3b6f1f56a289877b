"""The C-ABI library loads and exports every symbol include/psmc_b200.h declares (no compute without a GPU);
host-only entry points (factorize, unpack_stats) behave as documented."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "psmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(psmc_b200_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_typed():
    from psmc_b200 import _lib
    lib = _lib.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libpsmc_b200.so does not export %s" % n
        assert n in _lib.SYMBOLS, "psmc_b200/_lib.py does not type %s" % n
    assert set(_lib.SYMBOLS) == set(names)
    assert lib.psmc_b200_version() == 100


def test_missing_library_fails_loudly(tmp_path):
    from psmc_b200 import _lib
    with pytest.raises(_lib.LibraryNotBuilt):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_no_cpu_fallback_without_device():
    import psmc_b200
    lib = psmc_b200.load_library()
    if lib.psmc_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(psmc_b200.Psmc200Error) as ei:
        psmc_b200.EStep([np.zeros(10, dtype=np.int8)], 23)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_factorize_and_structure_check(oracle):
    import psmc_b200
    from helpers import make_model
    m = make_model(oracle, 64, seed=4)
    f = psmc_b200.factorize(m["a"])
    mod = psmc_b200.Model(m["a0"], m["e"], **f)
    assert np.max(np.abs(mod.dense() / m["a"] - 1)) < 1e-14
    capped = m["a"].copy()
    capped[:, 20] = capped[:, 20:].sum(axis=1); capped[:, 21:] = 0     # aux.c:115-127
    with pytest.raises(psmc_b200.Psmc200Error) as ei:
        psmc_b200.factorize(capped)
    assert ei.value.code == -4


def test_unpack_stats_adds_tiny_like_the_reference(oracle):
    """hmm_expect starts every count at HMM_TINY per sequence (khmm.c:305-308); the raw device vector does not"""
    from psmc_b200 import _lib
    lib = _lib.load_library()
    N, nseq = 5, 3
    raw = np.zeros(7 * N + 1)
    raw[0] = -12.5
    E = np.zeros((2, N)); v = [np.zeros(N) for _ in range(5)]
    st = _lib.CStats()
    dp = C.POINTER(C.c_double)
    st.E = E.ctypes.data_as(dp)
    for k, a in zip(("RL", "CL", "RU", "CU", "AD"), v):
        setattr(st, k, a.ctypes.data_as(dp))
    assert lib.psmc_b200_unpack_stats(N, raw.ctypes.data_as(dp), nseq, C.byref(st)) == 0
    tiny = nseq * 1e-25
    assert st.LL == -12.5
    assert np.allclose(E, tiny, rtol=1e-12, atol=0)
    k = np.arange(N)
    assert np.allclose(v[0], tiny * k, rtol=1e-12, atol=0) and np.allclose(v[3], tiny * k, rtol=1e-12, atol=0)      # RL, CU
    assert np.allclose(v[1], tiny * (N - 1 - k), rtol=1e-12, atol=0) and np.allclose(v[2], tiny * (N - 1 - k), rtol=1e-12, atol=0)
    assert np.allclose(v[4], tiny, rtol=1e-12, atol=0)
    raw[3] = np.nan
    assert lib.psmc_b200_unpack_stats(N, raw.ctypes.data_as(dp), nseq, C.byref(st)) == -5
