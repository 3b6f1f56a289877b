"""GPU parity of the E-step (include/psmc_b200.h) against the CPU oracle, through the C ABI.

Bar: FP64 relative 1e-10 on LL and on every expected count (north_star asks 1e-6 on the printed
LK/TR/RS values; the E-step itself is held to a much tighter bound so EM trajectories stay together)."""
import numpy as np
import pytest

from helpers import compare_stats, make_model, oracle_stats

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _model(m):
    from psmc_b200 import Model
    return Model.from_dense(m["a0"], m["a"], m["e"])


def _seqs(m, lengths, seed):
    from psmc_b200 import synth
    return synth.simulate_genome(m["a0"], m["a"], m["e"], lengths, seed, miss_frac=0.03, miss_mean=20)


@pytest.mark.parametrize("N", [5, 23, 33, 64, 100])
@pytest.mark.parametrize("chunk_len", [7, 64, 1 << 20])
def test_estep_matches_oracle_ragged(oracle, N, chunk_len):
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=N)
    seqs = _seqs(m, [1, 2, 3, 17, 300, 1000, 2049, 64, 128], seed=100 + N)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=chunk_len) as es:
        got = es.run(_model(m))
        info = es.info()
    assert info["n_seqs"] == len(seqs)
    compare_stats(got, want, TOL, N)


def test_generation1_kernels_still_match(oracle, monkeypatch):
    """PSMC_B200_GEN=1: the Kogge-Stone kernels (the backward fallback at 128 padded states) through the fast path"""
    from psmc_b200 import EStep
    monkeypatch.setenv("PSMC_B200_GEN", "1")
    N = 64
    m = make_model(oracle, N, seed=13)
    seqs = _seqs(m, [40000, 9000, 77], seed=14)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=2000) as es:
        es.set_warm(3000)
        got = es.run(_model(m))
    compare_stats(got, want, TOL, N)


def test_generations_agree_at_scale(oracle, monkeypatch):
    """3 M bins (far beyond what the oracle does in seconds): the two independently written kernel generations, with
    different chunk plans, overlaps and scan networks, must agree on LL and on every expected count"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=17)
    seqs = _seqs(m, [1500000, 900000, 600000, 333], seed=18)
    res = {}
    for gen in ("1", "2"):
        monkeypatch.setenv("PSMC_B200_GEN", gen)
        with EStep(seqs, N) as es:
            res[gen] = es.run(_model(m))
            info = es.info()
        assert info["fallbacks"] == 0 and info["fwd_mismatch"] < 1e-12 and info["bwd_mismatch"] < 1e-12
    compare_stats(res["2"], res["1"], TOL, N)


def test_chunking_is_exact(oracle):
    """the chunk plan must not change the result (it does for the reference's splitfa, which cuts the likelihood)"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=3)
    seqs = _seqs(m, [5000, 1234], seed=9)
    res = []
    for cl in (1 << 20, 1000, 97, 16, 5):
        with EStep(seqs, N, chunk_len=cl) as es:
            res.append(es.run(_model(m)))
    for r in res[1:]:
        compare_stats(r, res[0], 1e-11, N)


def test_estep_dense_entry_and_structure_check(oracle):
    from psmc_b200 import EStep, Psmc200Error
    N = 23
    m = make_model(oracle, N, seed=1)
    seqs = _seqs(m, [700, 300], seed=2)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=128) as es:
        got = es.run_dense(m["a0"], m["a"], m["e"])
        compare_stats(got, want, TOL, N)
        capped = m["a"].copy()          # psmc_cap_matrix (aux.c:115-127) destroys the rank structure
        capped[:, 10] = capped[:, 10:].sum(axis=1); capped[:, 11:] = 0
        with pytest.raises(Psmc200Error) as ei:
            es.run_dense(m["a0"], capped, m["e"])
        assert ei.value.code == -4


def test_dense_counts_match_oracle(oracle):
    """option: hmm_expect's dense A[N][N] (khmm.c:305-316) from the spilled backward rows and a tall-skinny product"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=61)
    seqs = _seqs(m, [30000, 7000, 1], seed=62)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=2000) as es:
        es.set_warm(3000)
        es.set_dense(True)
        got = es.run(_model(m))
        A = es.dense_counts()
    compare_stats(got, want, TOL, N)
    scale = np.maximum(np.abs(want["A"]), 1e-9 * np.abs(want["A"]).max())
    assert np.max(np.abs(A - want["A"]) / scale) < TOL


def test_missing_only_and_all_hom(oracle):
    from psmc_b200 import EStep
    N = 23
    m = make_model(oracle, N, seed=5)
    seqs = [np.full(500, 2, dtype=np.int8), np.zeros(700, dtype=np.int8), np.ones(40, dtype=np.int8)]
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=50) as es:
        got = es.run(_model(m))
    compare_stats(got, want, TOL, N)


def test_empty_sequences_are_skipped(oracle):
    from psmc_b200 import EStep
    N = 23
    m = make_model(oracle, N, seed=6)
    seqs = _seqs(m, [400], seed=1)
    want = oracle_stats(oracle, m, seqs)
    with EStep([np.zeros(0, dtype=np.int8), seqs[0], np.zeros(0, dtype=np.int8)], N, chunk_len=100) as es:
        got = es.run(_model(m))
        assert es.info()["n_seqs"] == 1
    compare_stats(got, want, TOL, N)


def test_medium_contig_auto_chunks(oracle):
    """200k bins, 64 states, automatic chunk plan (the oracle needs a few seconds for this)"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=11)
    seqs = _seqs(m, [150000, 50000], seed=12)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N) as es:
        got = es.run(_model(m))
        info = es.info()
    assert info["n_chunks"] >= 2
    compare_stats(got, want, TOL, N)


@pytest.mark.parametrize("warm", [0, 8, 300, 3000, 100000])
def test_warmup_certificate_and_fallback(oracle, warm):
    """fast path = warm-up overlaps + boundary certificate; a too-short overlap must be caught by the
    certificate and redone with the exact transfer-matrix path -- the result never depends on `warm`"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=31)
    seqs = _seqs(m, [60000, 20000, 500], seed=32)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=2500) as es:
        es.set_warm(warm)
        got = es.run(_model(m))
        info = es.info()
        got2 = es.run(_model(m))     # after a fallback the overlap has been doubled
        info2 = es.info()
    compare_stats(got, want, TOL, N)
    compare_stats(got2, want, TOL, N)
    assert info["fallbacks"] == 0 and info2["fallbacks"] == 0
    if warm == 0:
        assert info["ms"][0] > 0                                     # transfer-matrix path
    else:
        assert info["ms"][0] < 0.05                                  # fast path: no transfer matrices
        assert info["fwd_mismatch"] < 1e-12 and info["bwd_mismatch"] < 1e-12   # final certificate
    if warm in (8, 300):
        assert info["repaired_fwd"] > 0 and info["repaired_bwd"] > 0   # short overlaps are caught and repaired locally
    if warm == 100000:
        assert info["repaired_fwd"] == 0 and info["repaired_bwd"] == 0


def test_repeatable_bitwise(oracle):
    """same inputs, same call history -> the same bits (fixed-order reductions, no atomics in the data path).  A context
    adapts its per-boundary overlaps from one E-step to the next, so consecutive E-steps of ONE context agree to the
    certificate's 1e-12 rather than bitwise; two contexts with the same history agree bitwise, E-step by E-step."""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=2)
    seqs = _seqs(m, [30000], seed=3)
    runs = []
    for rep in range(2):
        with EStep(seqs, N, chunk_len=1000) as es:
            runs.append([es.run(_model(m)) for _ in range(3)])
    for a, b in zip(*runs):
        assert a["LL"] == b["LL"]
        for k in ("E", "RL", "CL", "RU", "CU", "AD"):
            assert np.array_equal(a[k], b[k])
    compare_stats(runs[0][2], runs[0][0], 1e-11, N)


@pytest.mark.parametrize("N", [23, 64])
def test_decode_matches_oracle(oracle, N):
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=21)
    seqs = _seqs(m, [3000, 555], seed=22)
    with EStep(seqs, N, chunk_len=200) as es:
        mod = _model(m)
        for i, s in enumerate(seqs):
            got = es.decode(mod if i == 0 else None, i, full=True, want_s=True)
            want = oracle.decode(m["a"], m["e"], m["a0"], s, full=True)
            f, b, sc = oracle.fwdbwd(m["a"], m["e"], m["a0"], s)
            assert np.max(np.abs(got["s"] / sc - 1)) < 1e-11
            assert np.max(np.abs(got["post"] - want["post"])) < 1e-11
            assert np.max(np.abs(got["p_recomb"] - want["p_recomb"])) < 1e-10
            assert np.max(np.abs(got["best_p"] - want["best_p"])) < 1e-11
            # argmax may legitimately differ only where two posteriors tie to ~1e-11
            diff = got["best_k"] != want["best_k"]
            if diff.any():
                srt = np.sort(want["post"][diff], axis=1)
                assert np.all(srt[:, -1] - srt[:, -2] < 1e-10)


# ------------------------------------------------------------------------------------------------
# Parity at the benchmark regime against the UNMODIFIED reference (oracle/_ref/libpsmcref.so = khmm.c compiled as is):
# default chunk plan (one resident wave, default overlaps, operator prediction on, default repair rounds), 64 states,
# two consecutive E-steps on one context (the second one runs with the boundary predictions of the first).
# ------------------------------------------------------------------------------------------------
def _bench_contig(L, seed):
    """a contig drawn exactly like bench.py's genome (true bottleneck history, 2 % missing data in runs)"""
    from psmc_b200 import host, synth
    n, nf, _ = host.parse_pattern("4+25*2+4+6")
    hm = host.model_from_params("4+25*2+4+6", np.concatenate([[0.05, 0.0125, 15.0], synth.bottleneck_lambdas(nf)]))
    return synth.simulate(hm["a0"], hm["model"].dense(), hm["e"], L, np.random.default_rng(seed))


def _ref_stats(ref, m, seqs):
    r = ref.estep(m["a"], m["e"], m["a0"], seqs)
    A = r["A"]          # hmm_exp_t::A summed over the records; its five structured marginals in numpy
    lo, up = np.tril(A, -1), np.triu(A, 1)
    r.update(RL=lo.sum(axis=1), CL=lo.sum(axis=0), RU=up.sum(axis=1), CU=up.sum(axis=0), AD=np.diag(A).copy())
    return r


@pytest.mark.parametrize("L,seed", [(500000, 20260925), (2489564, 20260926)])
def test_benchmark_regime_matches_unmodified_reference(ref, L, seed):
    from psmc_b200 import EStep
    N = 64
    seq = _bench_contig(L, seed)
    # EM's starting model (flat history, theta from the data: core.c:39) and a perturbed bottleneck model: the second
    # E-step sees a different model than the one its boundary predictions were made with, as in EM
    th = -np.log(1.0 - float((seq == 1).sum()) / float((seq < 2).sum()))
    m0 = ref.update_hmm("4+25*2+4+6", np.concatenate([[th, th / 5.0, 15.0], np.ones(28)]))
    m1 = make_model(ref, N, seed=7, theta=0.05)
    with EStep([seq], N) as es:
        for it, m in enumerate((m0, m1, m1)):
            got = es.run(_model(m))
            info = es.info()
            want = _ref_stats(ref, m, [seq])
            errs = compare_stats(got, want, TOL, N)
            print("L=%d E-step %d: chunks %d x %d, planned %d (mean overlaps %.0f / %.0f), failed fwd/bwd %d/%d, fallbacks %d, worst rel err %s"
                  % (L, it, info["n_chunks"], info["chunk_len"], info["planned"], info["avg_overlap_fwd"], info["avg_overlap_bwd"],
                     info["failed_fwd"], info["failed_bwd"], info["fallbacks"], {k: "%.1e" % v for k, v in errs.items()}))
            assert info["fallbacks"] == 0
            assert info["warm_len"] > 0                                 # the fast path (overlaps + certificate), default settings
            if it == 0:
                assert info["planned"] == 0 and info["n_chunks"] > 900  # first E-step: fixed overlaps, one resident wave of chunks
            else:
                assert info["planned"] == 1                             # then the overlaps the mixing probe of E-step 0 asked for
            assert info["fwd_mismatch"] < 1e-12 and info["bwd_mismatch"] < 1e-12


def test_decode_counts_records_as_given(oracle):
    """seq_id of psmc_b200_decode indexes the records as given to create: an empty record in front must not shift it"""
    from psmc_b200 import EStep, Psmc200Error
    N = 23
    m = make_model(oracle, N, seed=23)
    a, b = _seqs(m, [3000, 4111], seed=24)
    with EStep([np.zeros(0, dtype=np.int8), a, b], N, chunk_len=500) as es:
        mod = _model(m)
        with pytest.raises(Psmc200Error):
            es.decode(mod, 0)
        for i, s in ((1, a), (2, b)):
            got = es.decode(mod, i, full=False)
            want = oracle.decode(m["a"], m["e"], m["a0"], s, full=False)
            assert len(got["best_p"]) == len(s) and np.max(np.abs(got["best_p"] - want["best_p"])) < 1e-11


def _runs_of(best_k, best_p):
    out, start = [], 0
    for u in range(1, len(best_k) + 1):
        if u == len(best_k) or best_k[u] != best_k[start]:
            out.append((start, u - start, int(best_k[start]), float(best_p[start:u].max())))
            start = u
    return out


@pytest.mark.parametrize("N,chunk_len,warm", [(23, 1000, 0), (64, 997, 3000), (64, 0, -1), (100, 640, 2000)])
def test_decode_all_runs_bins_and_posteriors(oracle, N, chunk_len, warm):
    """whole-context decoding on the fast path (psmc_b200_decode_run): runs compacted on the device (-d), uint8 / float
    per-bin outputs and float posterior rows (-D), against the oracle's double-precision decode"""
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=121)
    a, b, c3 = _seqs(m, [60000, 1, 7333], seed=122)
    seqs = [a, np.zeros(0, dtype=np.int8), b, c3]
    with EStep(seqs, N, chunk_len=chunk_len) as es:
        if warm >= 0:
            es.set_warm(warm)
        got = es.decode_all(_model(m), runs=True, bins=True, post=True)
        info = es.info()
    assert info["fallbacks"] == 0
    runs = got["runs"]
    for i, s in enumerate(seqs):
        if len(s) == 0:
            assert got["seqs"][i] is None and not (runs["seq"] == i).any()
            continue
        want = oracle.decode(m["a"], m["e"], m["a0"], s, full=True)
        g = got["seqs"][i]
        assert np.max(np.abs(g["post"] - want["post"])) < 2e-7
        assert np.max(np.abs(g["p_recomb"] - want["p_recomb"])) < 1e-10
        assert np.max(np.abs(g["best_p"] - want["best_p"])) < 2e-7
        diff = g["best_k"] != want["best_k"]
        if diff.any():
            srt = np.sort(want["post"][diff], axis=1)
            assert np.all(srt[:, -1] - srt[:, -2] < 1e-10)
        sel = runs["seq"] == i
        mine = list(zip(runs["start"][sel], runs["len"][sel], runs["state"][sel], runs["max_p"][sel]))
        ref_runs = _runs_of(g["best_k"], want["best_p"] if not diff.any() else g["best_p"].astype(np.float64))
        assert [(x[0], x[1], x[2]) for x in mine] == [(x[0], x[1], x[2]) for x in ref_runs]
        assert max(abs(x[3] - y[3]) for x, y in zip(mine, ref_runs)) < (1e-10 if not diff.any() else 2e-7)


@pytest.mark.parametrize("chunk_len", [1500, 0])
def test_batch_of_models_matches_oracle_and_single_runs(oracle, chunk_len):
    """psmc_b200_set_batch: several models, each over its own multiset of the resident records, in one launch sequence"""
    from psmc_b200 import EStep
    N = 64
    ms = [make_model(oracle, N, seed=81 + r) for r in range(4)]
    seqs = _seqs(ms[0], [30000, 14000, 1, 22000, 9000], seed=82)
    rng = np.random.default_rng(5)
    mults = rng.integers(0, 3, size=(4, 5)).astype(np.int32)
    mults[:, 0] = [1, 0, 2, 1]
    with EStep(seqs, N, chunk_len=chunk_len) as es:
        es.set_warm(3000)
        es.set_batch(mults)
        got = es.run_batch([_model(m) for m in ms])
        got2 = es.run_batch([_model(m) for m in ms])
        assert es.info()["fallbacks"] == 0 and es.info()["n_models"] == 4
        alone = []
        for r in range(4):
            es.set_multiplicity(mults[r])
            alone.append(es.run(_model(ms[r])))
    for r in range(4):
        expanded = [s for s, k in zip(seqs, mults[r]) for _ in range(k)]
        want = oracle_stats(oracle, ms[r], expanded)
        compare_stats(got[r], want, TOL, N)
        compare_stats(got2[r], want, TOL, N)
        if chunk_len > 0:
            assert got[r]["LL"] == alone[r]["LL"]
            for k in ("E", "RL", "CL", "RU", "CU", "AD"):
                assert np.array_equal(got[r][k], alone[r][k]), k
        else:
            compare_stats(got[r], alone[r], 1e-11, N)


def test_staged_backward_equals_register_ring(oracle, monkeypatch):
    """bulk-asynchronous staging of the forward spill (cp.async.bulk + mbarrier, the default) vs the register prefetch ring:
    the same arithmetic, hence the same bits -- at a size where tiles wrap the ring many times"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=151)
    seqs = _seqs(m, [1, 2, 7, 8, 9, 70000, 10333, 64], seed=152)
    res = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("PSMC_B200_TMA", tma)
        with EStep(seqs, N, chunk_len=613) as es:
            es.set_warm(3000)
            res[tma] = es.run(_model(m))
    compare_stats(res["1"], oracle_stats(oracle, m, seqs), TOL, N)
    assert res["0"]["LL"] == res["1"]["LL"]
    for k in ("E", "RL", "CL", "RU", "CU", "AD"):
        assert np.array_equal(res["0"][k], res["1"][k]), k
