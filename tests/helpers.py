"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import numpy as np

from psmc_b200 import synth

PATTERNS = {23: "4+5*3+4", 64: "4+25*2+4+6", 5: "2+3", 33: "11*3", 100: "4+46*2+4"}


def make_model(oracle, N, seed=0, theta=0.07, ratio=5.0, max_t=15.0, flat=False):
    pat = PATTERNS[N]
    n, nf, pm = oracle.pattern(pat)
    rng = np.random.default_rng(seed)
    lam = np.ones(nf) if flat else synth.bottleneck_lambdas(nf) * np.exp(rng.normal(0, 0.2, nf))
    params = np.concatenate([[theta, theta / ratio, max_t], lam])
    m = oracle.update_hmm(pat, params)
    m["pattern"] = pat
    m["params"] = params
    return m


def rel_err(x, y, floor=1e-300):
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    return float(np.max(np.abs(x - y) / np.maximum(np.abs(y), floor))) if x.size else 0.0


def oracle_stats(oracle, m, seqs):
    r = oracle.estep(m["a"], m["e"], m["a0"], seqs)
    r.update(oracle.struct_stats(r["A"]))
    return r


def compare_stats(got, want, tol, N):
    """relative on LL; on the count vectors relative to the largest entry of each vector plus relative per entry"""
    errs = {"LL": abs(got["LL"] - want["LL"]) / max(abs(want["LL"]), 1e-300)}
    for k in ("E", "RL", "CL", "RU", "CU", "AD"):
        g = np.asarray(got[k]).ravel(); w = np.asarray(want[k]).ravel()
        scale = np.maximum(np.abs(w), 1e-9 * np.abs(w).max() + 1e-300)
        errs[k] = float(np.max(np.abs(g - w) / scale))
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, "parity failure (tol %g): %s" % (tol, errs)
    return errs
