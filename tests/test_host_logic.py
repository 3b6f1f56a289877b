"""Host driver logic (host/*.c through psmc_b200/host.py) against the oracle and the golden files: pattern parser,
params -> factored model, O(N) objective + Hooke-Jeeves M-step, .psmcfa reader, bootstrap resampler."""
import gzip
import os
import re

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("pat,n,nf", [("4+5*3+4", 22, 7), ("4+25*2+4+6", 63, 28), ("1+1", 1, 2), ("3*2+1", 6, 4), ("64*1", 63, 64)])
def test_pattern_parser(oracle, pat, n, nf):
    from psmc_b200 import host
    got = host.parse_pattern(pat)
    assert got[0] == n and got[1] == nf
    assert np.array_equal(got[2], oracle.pattern(pat)[2])


def test_bad_pattern_is_rejected():
    from psmc_b200 import host
    with pytest.raises(ValueError):
        host.parse_pattern("4+x")


@pytest.mark.parametrize("pat", ["4+5*3+4", "4+25*2+4+6", "1+1"])
def test_model_matches_oracle(oracle, pat):
    from psmc_b200 import host
    n, nf, _ = oracle.pattern(pat)
    rng = np.random.default_rng(nf)
    for trial in range(5):
        params = np.concatenate([[10 ** rng.uniform(-2.5, -1), 10 ** rng.uniform(-3, -1.5), rng.uniform(5, 20)], np.exp(rng.normal(0, 0.5, nf))])
        m, mo = host.model_from_params(pat, params), oracle.update_hmm(pat, params)
        for k in ("t", "sigma", "e"):                         # same operations in the same order: bit-identical
            assert np.array_equal(m[k], mo[k]), k
        assert m["C_pi"] == mo["C_pi"] and m["C_sigma"] == mo["C_sigma"]
        assert np.max(np.abs(m["model"].dense() / mo["a"] - 1)) < 5e-16    # factors vs dense entries: 1-2 ulp
        assert np.array_equal(host.avg_t(pat, params), oracle.avg_t(pat, params, mo))


def test_hooke_jeeves_same_path_as_oracle(oracle):
    from psmc_b200 import host
    f = lambda x: float(((x - np.arange(len(x))) ** 2).sum() + np.abs(x).sum() + np.sin(3 * x).sum())  # noqa: E731
    a, b = host.hooke_jeeves(f, np.ones(6)), oracle.hj(f, np.ones(6))
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2]


def test_structured_objective_equals_dense_objective(oracle):
    """O(N) objective on the marginals == hmm_Q on the dense counts (khmm.c:363-382) to rounding"""
    from psmc_b200 import host
    d = np.load(os.path.join(G, "estep_64.npz"))
    pat = str(d["pattern"])
    res = host.mstep(pat, d["params"], d["E"], A=d["A"])
    q0 = oracle.Q0(d["A"], d["E"])
    assert abs(res["Q0_offset"] - q0) <= 1e-12 * abs(q0)
    qd = oracle.Q(d["a"], d["e"], d["A"], d["E"], q0)
    assert abs(res["Q0"] - qd) <= 1e-9 * abs(qd) + 1e-9
    assert res["Q1"] >= res["Q0"] and res["calls"] > 100


def test_mstep_within_the_reference_noise_floor(oracle):
    """The reference's own M-step moves by ~1e-6..1e-4 under a ONE-ULP perturbation of the counts (Hooke-Jeeves takes
    a different path, DESIGN.md 'Parity'): the O(N) M-step must stay inside that band and reach the same optimum value."""
    from psmc_b200 import host
    d = np.load(os.path.join(G, "estep_23.npz"))
    pat = str(d["pattern"])
    A, E = d["A"], d["E"]

    def dense_mstep(A_, E_):
        q0 = oracle.Q0(A_, E_)
        last = [None]

        def func(x):
            p = np.abs(x); last[0] = p.copy()
            mm = oracle.update_hmm(pat, p)
            return -oracle.Q(mm["a"], mm["e"], A_, E_, q0)
        fx, _, calls = oracle.hj(func, d["params"])
        return -fx, last[0]
    q_ref, p_ref = dense_mstep(A, E)
    rng = np.random.default_rng(1)
    q_pert, p_pert = dense_mstep(A * (1 + 2.2e-16 * rng.standard_normal(A.shape)), E)
    noise = np.max(np.abs(p_pert / p_ref - 1))
    res = host.mstep(pat, d["params"], E, A=A)
    dev = np.max(np.abs(res["params"] / p_ref - 1))
    assert abs(res["Q1"] - q_ref) <= 1e-8 * abs(q_ref)
    assert dev < max(50 * noise, 5e-4), (dev, noise)


def test_psmcfa_reader_matches_reference_header():
    """n_seqs / sum_L / sum_n printed by the reference for the golden inputs (cli.c:224)"""
    from psmc_b200 import host, psmcfa
    for stem in ("c1", "small64"):
        hdr = open(os.path.join(G, stem + ".psmc")).read()
        m = re.search(r"MM\tn_seqs:(\d+), sum_L:(\d+), sum_n:(\d+)", hdr)
        names, seqs, sum_L, sum_n = host.read_psmcfa(os.path.join(G, stem + ".psmcfa.gz"))
        assert (len(seqs), sum_L, sum_n) == tuple(int(x) for x in m.groups())
        names2, seqs2 = psmcfa.read_psmcfa(os.path.join(G, stem + ".psmcfa.gz"))
        assert names == names2 and all(np.array_equal(a, b) for a, b in zip(seqs, seqs2))


def test_psmcfa_reader_grammar(tmp_path):
    """conv_table semantics (cli.c:15-32), FASTQ records, multi-line bodies, comments in headers (kseq.h:173-217)"""
    from psmc_b200 import host
    p = tmp_path / "x.fa"
    p.write_text(">s1 some comment\nTTKkN\nacgt01\nMRSWYmrswy\nXZ-*\n@q1\nTKTK\n+\nIIII\n>empty\n>s3\nT\n")
    names, seqs, sum_L, sum_n = host.read_psmcfa(str(p))
    assert names == ["s1", "q1", "empty", "s3"]
    assert seqs[0].tolist() == [0, 0, 1, 1, 2] + [0, 0, 0, 0, 0, 1] + [1] * 10 + [2, 2, 2, 2]
    assert seqs[1].tolist() == [0, 1, 0, 1]
    assert len(seqs[2]) == 0 and seqs[3].tolist() == [0]
    assert sum_n == 2 + 1 + 10 + 2 and sum_L == 20 + 4 + 1
    gz = tmp_path / "x.fa.gz"
    with gzip.open(gz, "wb") as fp:
        fp.write(p.read_bytes())
    assert host.read_psmcfa(str(gz))[0] == names


def test_resampler_properties(tmp_path):
    """aux.c:8-47: whole records drawn with replacement until the total length is as close as possible to the original"""
    import ctypes as C
    from psmc_b200 import host
    L = host.load_host()
    lens = [500, 700, 300, 900, 100]
    p = tmp_path / "s.fa"
    p.write_text("".join(">r%d\n%s\n" % (i, "TK" * (n // 2)) for i, n in enumerate(lens)))
    rng = np.random.default_rng(3)
    RND = C.CFUNCTYPE(C.c_double)
    for trial in range(20):
        h = L.psmch_py_read(str(p).encode())
        cb = RND(lambda: float(rng.random()))
        L.psmch_py_resample.argtypes = [C.c_void_p, RND]
        L.psmch_py_resample(h, cb)
        n = L.psmch_py_read_n(h)
        got = [L.psmch_py_read_len(h, i) for i in range(n)]
        assert all(x in lens for x in got)
        assert abs(sum(got) - sum(lens)) <= max(lens)
        assert L.psmch_py_read_sum(h, 0) == sum(got)          # all bins are informative in this input
        L.psmch_py_read_free(h)


@pytest.mark.parametrize("npz", ["estep_23.npz", "estep_64.npz"])
def test_speculative_mstep_is_the_sequential_search(npz, monkeypatch):
    """host/spec.c: the -step point of every Hooke-Jeeves probe is evaluated on a helper thread while the caller evaluates the
    +step point.  Same search path, same counted calls, same last evaluated point, same optimum -- bit for bit."""
    import time
    from psmc_b200 import host
    d = np.load(os.path.join(G, npz))
    pat = str(d["pattern"])
    out = {}
    for spec in ("0", "1", "3", "5", "7"):
        monkeypatch.setenv("PSMC_B200_MSTEP_SPEC", spec)
        t0 = time.perf_counter()
        out[spec] = host.mstep(pat, d["params"], d["E"], A=d["A"])
        out[spec]["t"] = time.perf_counter() - t0
    for k in ("1", "3", "5", "7"):
        a, b = out["0"], out[k]
        assert a["calls"] == b["calls"] and a["Q1"] == b["Q1"]
        assert np.array_equal(a["params"], b["params"])
