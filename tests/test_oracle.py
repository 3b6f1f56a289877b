"""The oracle (oracle/psmc_oracle.c) is pinned: (1) bit-exact against the committed golden vectors that the
UNMODIFIED reference produced (tools/make_golden.py), (2) bit-exact against the reference itself when oracle/_ref
is present (dev container and, prebuilt, the GPU box)."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(tag):
    d = np.load(os.path.join(G, "estep_%s.npz" % tag))
    lens = d["lens"]
    cat = d["seqs"].astype(np.int8)
    seqs, p = [], 0
    for L in lens:
        seqs.append(cat[p:p + L]); p += L
    return d, seqs


@pytest.mark.parametrize("tag", ["23", "64"])
def test_oracle_matches_golden_bit_exact(oracle, tag):
    d, seqs = _golden(tag)
    pat = str(d["pattern"])
    m = oracle.update_hmm(pat, d["params"])
    for k in ("a", "e", "a0", "sigma"):
        assert np.array_equal(m[k], d[k]), k
    assert np.array_equal(m["t"][:-1], d["t"][:-1])
    assert m["C_pi"] == float(d["C_pi"]) and m["C_sigma"] == float(d["C_sigma"])
    r = oracle.estep(d["a"], d["e"], d["a0"], seqs)
    assert r["LL"] == float(d["LL"])
    assert np.array_equal(r["A"], d["A"]) and np.array_equal(r["E"], d["E"])
    assert oracle.Q0(r["A"], r["E"]) == float(d["Q0"])


@pytest.mark.parametrize("pat", ["4+5*3+4", "4+25*2+4+6", "1+1", "3*2+1", "2*1+3*2"])
def test_oracle_matches_reference_bit_exact(oracle, ref, pat):
    n, nf, pm = oracle.pattern(pat)
    n2, nf2, pm2 = ref.pattern(pat)
    assert (n, nf) == (n2, nf2) and np.array_equal(pm, pm2)
    rng = np.random.default_rng(n)
    params = np.concatenate([[10 ** rng.uniform(-2.5, -1), 10 ** rng.uniform(-3, -1.5), rng.uniform(5, 20)], np.exp(rng.normal(0, 0.4, nf))])
    mo, mr = oracle.update_hmm(pat, params), ref.update_hmm(pat, params)
    for k in ("a", "e", "a0", "sigma"):
        assert np.array_equal(mo[k], mr[k]), k
    seqs = [rng.choice(np.array([0, 1, 2], dtype=np.int8), size=L, p=[0.9, 0.07, 0.03]) for L in (400, 1, 2, 33)]
    eo, er = oracle.estep(mo["a"], mo["e"], mo["a0"], seqs), ref.estep(mo["a"], mo["e"], mo["a0"], seqs)
    assert eo["LL"] == er["LL"]
    for k in ("A", "E", "A0"):
        assert np.array_equal(eo[k], er[k]), k
    q0 = oracle.Q0(eo["A"], eo["E"])
    qr, q0r = ref.Q_ref(mo["a"], mo["e"], mo["a0"], eo["A"], eo["E"])
    assert q0 == q0r and oracle.Q(mo["a"], mo["e"], eo["A"], eo["E"], q0) == qr
    fo, fr = oracle.fwdbwd(mo["a"], mo["e"], mo["a0"], seqs[0]), ref.fwdbwd(mo["a"], mo["e"], mo["a0"], seqs[0])
    assert all(np.array_equal(x, y) for x, y in zip(fo, fr))
    do, dr = oracle.decode(mo["a"], mo["e"], mo["a0"], seqs[0]), ref.decode(mo["a"], mo["e"], mo["a0"], seqs[0])
    assert all(np.array_equal(do[k], dr[k]) for k in do)


def test_hooke_jeeves_same_path_as_reference(oracle, ref):
    f = lambda x: float(((x - np.arange(len(x))) ** 2).sum() + np.abs(x).sum() + np.sin(3 * x).sum())  # noqa: E731
    a, b = oracle.hj(f, np.ones(6)), ref.hj(f, np.ones(6))
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_structured_view_of_dense_counts(oracle):
    """marginals RL/CL/RU/CU/AD and the factor extraction used by the parity tests"""
    d, seqs = _golden("64")
    S = oracle.struct_stats(d["A"])
    A = d["A"]; N = A.shape[0]
    lo, up = np.tril(A, -1), np.triu(A, 1)
    assert np.allclose(S["RL"], lo.sum(1), rtol=1e-13) and np.allclose(S["CL"], lo.sum(0), rtol=1e-13)
    assert np.allclose(S["RU"], up.sum(1), rtol=1e-13) and np.allclose(S["CU"], up.sum(0), rtol=1e-13)
    assert np.array_equal(S["AD"], A.diagonal())
    F = oracle.factors(d["a"])
    k, l = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    rec = np.where(l < k, np.outer(F["U"], F["V"]), np.where(l > k, np.outer(F["W"], F["Z"]), np.diag(F["D"])))
    assert np.max(np.abs(rec / d["a"] - 1)) < 1e-14          # diagonal + rank-1 lower + rank-1 upper
