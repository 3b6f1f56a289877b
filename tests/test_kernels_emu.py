"""The CUDA kernels' SOURCE executed on the host by a SIMT emulation (tests/emu), checked against the oracle.

Why: the GPU is a scarce resource and lives elsewhere; these tests run everywhere and catch logic errors in the kernels
(lane-group scans, chunk plans, warm-up / certificate / repair, software-pipelined stores) before GPU time is spent.
tests/emu/libpsmc_b200_emu.so is psmc_b200/csrc/psmc_estep.cu itself, compiled by g++ with CUDA threads as fibers.  It is
test infrastructure: nothing under psmc_b200/ or host/ loads it, and the `-m gpu` tests remain the parity tests proper.
Sizes are small (the emulation is slow); tolerance is the same 1e-10 as on the GPU."""
import os
import subprocess

import numpy as np
import pytest

from helpers import compare_stats, make_model, oracle_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
TOL = 1e-10


@pytest.fixture(scope="module")
def emu_built():
    subprocess.run(["make", "-C", EMU_DIR], check=True, capture_output=True)
    return EMU_DIR


@pytest.fixture(params=["libpsmc_b200_emu.so", "libpsmc_b200_emu_boost.so"], ids=["emu", "emu-boost"])
def emu(request, emu_built, monkeypatch):
    """the ctypes mirror (psmc_b200.EStep ...) bound to the emulated library for the duration of one test"""
    from psmc_b200 import _lib
    lib = _lib.load_library(path=os.path.join(emu_built, request.param))
    monkeypatch.setattr(_lib, "_lib", lib)
    monkeypatch.setenv("PSMC_EMU_SMS", "2")
    return lib


@pytest.fixture
def emu_plain(emu_built, monkeypatch):
    from psmc_b200 import _lib
    lib = _lib.load_library(path=os.path.join(emu_built, "libpsmc_b200_emu.so"))
    monkeypatch.setattr(_lib, "_lib", lib)
    return lib


def _model(m):
    from psmc_b200 import Model
    return Model.from_dense(m["a0"], m["a"], m["e"])


def _seqs(m, lengths, seed):
    from psmc_b200 import synth
    return synth.simulate_genome(m["a0"], m["a"], m["e"], lengths, seed, miss_frac=0.03, miss_mean=20)


@pytest.mark.parametrize("gen", ["1", "2"])
@pytest.mark.parametrize("N,chunk_len", [(5, 7), (23, 64), (33, 16), (64, 7), (64, 1 << 20), (100, 64)])
def test_emulated_estep_matches_oracle_ragged(oracle, emu, monkeypatch, N, chunk_len, gen):
    """both kernel generations, every lane-group layout (NP = 32 / 64 / 128), ragged records, transfer-matrix path"""
    from psmc_b200 import EStep
    monkeypatch.setenv("PSMC_B200_GEN", gen)
    monkeypatch.setenv("PSMC_B200_WARM", "0")
    m = make_model(oracle, N, seed=N)
    seqs = _seqs(m, [1, 2, 3, 17, 150, 260, 64], seed=100 + N)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=chunk_len) as es:
        got = es.run(_model(m))
    compare_stats(got, want, TOL, N)


@pytest.mark.parametrize("N", [23, 64, 100])
@pytest.mark.parametrize("warm", [8, 60, 2500])
def test_emulated_warmup_certificate_and_repair(oracle, emu, N, warm):
    """fast path: warm-up overlaps + certificate; short overlaps are caught and repaired, long ones pass untouched"""
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=31)
    seqs = _seqs(m, [3000, 1100, 40], seed=32)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=500) as es:
        es.set_warm(warm)
        got = es.run(_model(m))
        info = es.info()
    compare_stats(got, want, TOL, N)
    assert info["fallbacks"] == 0
    assert info["fwd_mismatch"] < 1e-12 and info["bwd_mismatch"] < 1e-12
    if warm == 8:
        assert info["repaired_fwd"] > 0 and info["repaired_bwd"] > 0
    if warm == 2500:
        assert info["repaired_fwd"] == 0 and info["repaired_bwd"] == 0


@pytest.mark.parametrize("N", [23, 64])
@pytest.mark.parametrize("warm", [0, 60])
def test_emulated_sixteen_lane_groups_for_forward_and_overlap_kernels(oracle, emu_plain, monkeypatch, N, warm):
    """PSMC_B200_G2_FWD / PSMC_B200_G2_BWW = 16: the forward, forward-repair and backward-overlap kernels with 16-lane groups
    (2 or 4 states per lane instead of 4 or 8) -- a tuning knob for small shards, same results"""
    from psmc_b200 import EStep
    monkeypatch.setenv("PSMC_B200_G2_FWD", "16")
    monkeypatch.setenv("PSMC_B200_G2_BWW", "16")
    m = make_model(oracle, N, seed=41)
    seqs = _seqs(m, [2000, 700, 33, 1], seed=42)
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=300) as es:
        es.set_warm(warm)
        got = es.run(_model(m))
        info = es.info()
    compare_stats(got, want, TOL, N)
    assert info["fallbacks"] == 0
    if warm:
        assert info["repaired_fwd"] > 0


def test_emulated_chunking_is_exact(oracle, emu_plain):
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=3)
    seqs = _seqs(m, [700, 234], seed=9)
    res = []
    for cl in (1 << 20, 97, 16, 5):
        with EStep(seqs, N, chunk_len=cl) as es:
            res.append(es.run(_model(m)))
    for r in res[1:]:
        compare_stats(r, res[0], 1e-11, N)


def test_emulated_edge_tracks(oracle, emu_plain):
    from psmc_b200 import EStep
    N = 23
    m = make_model(oracle, N, seed=5)
    seqs = [np.full(200, 2, dtype=np.int8), np.zeros(300, dtype=np.int8), np.ones(40, dtype=np.int8)]
    want = oracle_stats(oracle, m, seqs)
    with EStep([np.zeros(0, dtype=np.int8)] + seqs, N, chunk_len=50) as es:
        got = es.run(_model(m))
        assert es.info()["n_seqs"] == 3
    compare_stats(got, want, TOL, N)


def test_emulated_multiplicity_and_repeatability(oracle, emu_plain):
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=8)
    seqs = _seqs(m, [300, 200, 250, 100], seed=4)
    mult = [2, 0, 1, 3]
    expanded = [s for s, k in zip(seqs, mult) for _ in range(k)]
    want = oracle_stats(oracle, m, expanded)
    with EStep(seqs, N, chunk_len=64) as es:
        es.set_multiplicity(mult)
        a = es.run(_model(m))
        b = es.run(_model(m))
    compare_stats(a, want, TOL, N)
    assert a["LL"] == b["LL"]
    for k in ("E", "RL", "CL", "RU", "CU", "AD"):
        assert np.array_equal(a[k], b[k])


@pytest.mark.parametrize("N", [23, 64])
def test_emulated_decode_matches_oracle(oracle, emu_plain, N):
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=21)
    seqs = _seqs(m, [600, 155], seed=22)
    with EStep(seqs, N, chunk_len=100) as es:
        mod = _model(m)
        for i, s in enumerate(seqs):
            got = es.decode(mod if i == 0 else None, i, full=True, want_s=True)
            want = oracle.decode(m["a"], m["e"], m["a0"], s, full=True)
            f, b, sc = oracle.fwdbwd(m["a"], m["e"], m["a0"], s)
            assert np.max(np.abs(got["s"] / sc - 1)) < 1e-11
            assert np.max(np.abs(got["post"] - want["post"])) < 1e-11
            assert np.max(np.abs(got["p_recomb"] - want["p_recomb"])) < 1e-10
            assert np.max(np.abs(got["best_p"] - want["best_p"])) < 1e-11


def test_emulated_decode_counts_records_as_given(oracle, emu_plain):
    """seq_id of psmc_b200_decode indexes the records as given to create: an empty record in front must not shift it"""
    from psmc_b200 import EStep, Psmc200Error
    N = 23
    m = make_model(oracle, N, seed=23)
    a, b = _seqs(m, [300, 411], seed=24)
    with EStep([np.zeros(0, dtype=np.int8), a, b], N, chunk_len=100) as es:
        mod = _model(m)
        with pytest.raises(Psmc200Error):
            es.decode(mod, 0)                      # empty record: nothing to decode
        for i, s in ((1, a), (2, b)):
            got = es.decode(mod, i, full=False)
            want = oracle.decode(m["a"], m["e"], m["a0"], s, full=False)
            assert len(got["best_p"]) == len(s)
            assert np.max(np.abs(got["best_p"] - want["best_p"])) < 1e-11
        with pytest.raises(Psmc200Error):
            es.decode(mod, 3)


@pytest.mark.parametrize("N,mult", [(23, None), (64, None), (64, [2, 0, 1])])
def test_emulated_dense_counts(oracle, emu_plain, N, mult):
    """option: hmm_expect's dense A[N][N] (khmm.c:305-316) from the spilled backward rows and a tall-skinny product"""
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=61)
    seqs = _seqs(m, [900, 340, 1], seed=62)
    expanded = seqs if mult is None else [s for s, k in zip(seqs, mult) for _ in range(k)]
    want = oracle_stats(oracle, m, expanded)
    with EStep(seqs, N, chunk_len=150) as es:
        es.set_warm(300)
        es.set_dense(True)
        if mult is not None:
            es.set_multiplicity(mult)
        got = es.run(_model(m))
        A = es.dense_counts()
    compare_stats(got, want, TOL, N)
    scale = np.maximum(np.abs(want["A"]), 1e-9 * np.abs(want["A"]).max())
    assert np.max(np.abs(A - want["A"]) / scale) < TOL


def test_emulated_dense_counts_refused_beyond_64_states(oracle, emu_plain):
    """the dense-count option needs the generation-2 backward kernels (at most 64 states): it must say so, not misbehave"""
    from psmc_b200 import EStep, Psmc200Error
    N = 100
    m = make_model(oracle, N, seed=71)
    seqs = _seqs(m, [200], seed=72)
    with EStep(seqs, N, chunk_len=64) as es:
        with pytest.raises(Psmc200Error):
            es.set_dense(True)
        with pytest.raises(Psmc200Error):
            es.dense_counts()
        got = es.run(_model(m))          # the E-step itself is unaffected
    compare_stats(got, oracle_stats(oracle, m, seqs), TOL, N)


def test_emulated_decode_128_padded_states(oracle, emu_plain):
    from psmc_b200 import EStep
    N = 100
    m = make_model(oracle, N, seed=73)
    seqs = _seqs(m, [400, 90], seed=74)
    with EStep(seqs, N, chunk_len=100) as es:
        got = es.decode(_model(m), 0, full=True, want_s=True)
    want = oracle.decode(m["a"], m["e"], m["a0"], seqs[0], full=True)
    assert np.max(np.abs(got["post"] - want["post"])) < 1e-11
    assert np.max(np.abs(got["best_p"] - want["best_p"])) < 1e-11


@pytest.mark.parametrize("chunk_len,warm", [(150, 0), (150, 400), (0, 400)])
def test_emulated_batch_of_models(oracle, emu_plain, chunk_len, warm):
    """psmc_b200_set_batch: three models, each over its own multiset of the resident records, in one launch sequence;
    every model must match the oracle on its expanded record list, and (fixed chunk length) match a run of its own bit for bit"""
    from psmc_b200 import EStep
    N = 23
    ms = [make_model(oracle, N, seed=81 + r) for r in range(3)]
    seqs = _seqs(ms[0], [700, 340, 1, 520], seed=82)
    mults = np.array([[1, 0, 2, 1], [0, 3, 0, 1], [2, 1, 1, 0]], dtype=np.int32)
    with EStep(seqs, N, chunk_len=chunk_len) as es:
        es.set_warm(warm)
        es.set_batch(mults)
        assert es.info()["n_models"] == 3
        got = es.run_batch([_model(m) for m in ms])
        got2 = es.run_batch([_model(m) for m in ms])          # a second E-step of the same batch (operator predictions on)
        info = es.info()
        assert info["fallbacks"] == 0
        alone = []
        for r in range(3):
            es.set_multiplicity(mults[r])                      # leaves batch mode
            alone.append(es.run(_model(ms[r])))
        with pytest.raises(Exception):
            es.run_batch([_model(m) for m in ms])              # 3 models on a single-model context
    for r in range(3):
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


def test_emulated_batch_rejects_bad_input(oracle, emu_plain):
    from psmc_b200 import EStep, Psmc200Error
    N = 23
    m = make_model(oracle, N, seed=91)
    seqs = _seqs(m, [300, 200], seed=92)
    with EStep(seqs, N, chunk_len=100) as es:
        with pytest.raises(Psmc200Error):
            es.set_batch(np.array([[1, -1]], dtype=np.int32))
        es.set_batch(np.array([[1, 1], [0, 2]], dtype=np.int32))
        with pytest.raises(Psmc200Error):
            es.run(_model(m))                                   # single-model call on a batch context
        with pytest.raises(Psmc200Error):
            es.decode(_model(m), 0)
        with pytest.raises(Psmc200Error):
            es.set_dense(True)
        es.set_multiplicity(None)
        got = es.run(_model(m))
    compare_stats(got, oracle_stats(oracle, m, seqs), TOL, N)


def test_emulated_deep_repair_cascade_stays_on_the_fast_path(oracle, emu_plain):
    """chunks much shorter than the mixing length: repairs cascade over many rounds.  The number of rounds adapts and a
    failed certificate is first retried on the fast path with more rounds; the exact transfer-operator fallback stays unused"""
    from psmc_b200 import EStep
    N = 64
    m = make_model(oracle, N, seed=33)
    seqs = _seqs(m, [5000, 900], seed=34)
    seqs[0][1500:3600] = 0            # a long homozygous tract: slow mixing across many short chunks
    want = oracle_stats(oracle, m, seqs)
    with EStep(seqs, N, chunk_len=40) as es:
        es.set_warm(60)
        rounds = []
        for it in range(3):
            got = es.run(_model(m))
            info = es.info()
            rounds.append((info["repair_rounds"], info["warm_redos"], info["fallbacks"]))
            compare_stats(got, want, TOL, N)
    assert rounds[-1][2] == 0, rounds             # never the exact fallback
    assert rounds[-1][1] == 1, rounds             # one retry on the fast path (first E-step), none afterwards
    assert rounds[-1][0] > 3, rounds              # the cascade was deeper than the default three rounds: adapted


def _runs_of(best_k, best_p):
    """reference semantics of the DC lines (aux.c:165-182): runs of the argmax state with their maximum posterior"""
    out = []
    start = 0
    for u in range(1, len(best_k) + 1):
        if u == len(best_k) or best_k[u] != best_k[start]:
            out.append((start, u - start, int(best_k[start]), float(best_p[start:u].max())))
            start = u
    return out


@pytest.mark.parametrize("N,chunk_len,warm", [(23, 100, 0), (64, 97, 250), (64, 1 << 20, 250), (100, 64, 150)])
def test_emulated_decode_all_runs_bins_and_posteriors(oracle, emu_plain, N, chunk_len, warm):
    """whole-context decoding on the generation-2 kernels and the fast path: runs compacted on the device (-d), per-bin
    uint8 / float outputs and full posterior rows (-D), against the oracle's double-precision decode"""
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=121)
    a, b, c3 = _seqs(m, [900, 1, 333], seed=122)
    seqs = [a, np.zeros(0, dtype=np.int8), b, c3]
    with EStep(seqs, N, chunk_len=chunk_len) as es:
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
        assert np.max(np.abs(g["post"] - want["post"])) < 2e-7                     # float rows
        assert np.max(np.abs(g["p_recomb"] - want["p_recomb"])) < 1e-10
        assert np.max(np.abs(g["best_p"] - want["best_p"])) < 2e-7
        diff = g["best_k"] != want["best_k"]
        if diff.any():                                                             # argmax may differ only at ties
            srt = np.sort(want["post"][diff], axis=1)
            assert np.all(srt[:, -1] - srt[:, -2] < 1e-10)
        sel = runs["seq"] == i
        mine = list(zip(runs["start"][sel], runs["len"][sel], runs["state"][sel], runs["max_p"][sel]))
        ref_runs = _runs_of(g["best_k"], want["best_p"] if not diff.any() else g["best_p"].astype(np.float64))
        assert [(x[0], x[1], x[2]) for x in mine] == [(x[0], x[1], x[2]) for x in ref_runs]
        tol = 1e-10 if not diff.any() else 2e-7
        assert max(abs(x[3] - y[3]) for x, y in zip(mine, ref_runs)) < tol
        assert sum(x[1] for x in mine) == len(s)


def test_emulated_mixing_probe_plans_the_overlaps(oracle, emu_plain, monkeypatch):
    """automatic chunk plan: E-step 0 runs with the fixed overlap and measures the local mixing rate (k_probe); from E-step 1
    on every boundary gets the overlap the probe asks for, chunk lengths are balanced, slow tracts are known in advance.
    Results stay exact (certificate) whatever the probe says."""
    from psmc_b200 import EStep
    monkeypatch.setenv("PSMC_EMU_SMS", "3")
    N = 64
    m = make_model(oracle, N, seed=141)
    seqs = _seqs(m, [26000, 9000, 300], seed=142)
    seqs[0][8000:14000] = 0              # a long homozygous tract: no admissible overlap gets through it
    want = oracle_stats(oracle, m, seqs)
    m2 = make_model(oracle, N, seed=143)
    want2 = oracle_stats(oracle, m2, seqs)
    with EStep(seqs, N) as es:
        es.set_warm(1500)
        infos = []
        for it in range(4):
            mm, ww = (m, want) if it != 2 else (m2, want2)     # the model moves between E-steps, the plan is one E-step old
            got = es.run(_model(mm))
            infos.append(es.info())
            compare_stats(got, ww, TOL, N)
    assert infos[0]["planned"] == 0 and infos[1]["planned"] == 1
    assert infos[1]["avg_overlap_fwd"] < 1500 and infos[1]["avg_overlap_bwd"] < 2000
    assert all(i["fallbacks"] == 0 for i in infos)
    assert infos[1]["slow_fwd"] > 0 and infos[1]["slow_bwd"] > 0      # the tract
    print([(i["planned"], i["n_chunks"], i["chunk_len"], round(i["avg_overlap_fwd"]), round(i["avg_overlap_bwd"]), i["slow_fwd"], i["slow_bwd"],
            i["failed_fwd"], i["failed_bwd"], i["repair_rounds"]) for i in infos])


@pytest.mark.parametrize("chunk_len", [17, 61, 200])   # 2, 7 and 25 tiles of 8 rows per chunk: all-edge, mixed, mostly the branch-free interior body
@pytest.mark.parametrize("N", [23, 64])
def test_emulated_staged_backward_equals_register_ring(oracle, emu_plain, monkeypatch, N, chunk_len):
    """the backward pass with the forward spill staged through shared memory (bulk-asynchronous copies + mbarrier, the
    default) performs the same arithmetic as the register-prefetch variant: same bits, ragged records, dense option included"""
    from psmc_b200 import EStep
    m = make_model(oracle, N, seed=151)
    seqs = _seqs(m, [1, 2, 7, 8, 9, 700, 1033, 64], seed=152)
    res = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("PSMC_B200_TMA", tma)
        with EStep(seqs, N, chunk_len=chunk_len) as es:
            es.set_warm(150)
            es.set_dense(True)
            res[tma] = (es.run(_model(m)), es.dense_counts())
    compare_stats(res["1"][0], oracle_stats(oracle, m, seqs), TOL, N)
    assert res["0"][0]["LL"] == res["1"][0]["LL"]
    for k in ("E", "RL", "CL", "RU", "CU", "AD"):
        assert np.array_equal(res["0"][0][k], res["1"][0][k]), k
    assert np.array_equal(res["0"][1], res["1"][1])
