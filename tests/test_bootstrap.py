"""Bootstrap pipeline: splitfa rule, replicate draw, multiplicity-weighted E-step, batch driver.

CPU part: host logic against the reference's own `splitfa` (oracle/_ref, built from the reference sources) and the copying
resampler.  GPU part: psmc_b200_set_multiplicity against the oracle on the expanded record list, and
`psmc --replicates R --seed S` against R separate `psmc -b --seed S+r` runs (what the reference's README recipe does)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from helpers import compare_stats, make_model, oracle_stats
from psmc_text import compare_rounds, parse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
PSMC = os.path.join(ROOT, "host", "psmc")
SPLITFA = os.path.join(ROOT, "oracle", "_ref", "splitfa")


def test_split_rule_matches_reference_splitfa(tmp_path):
    from psmc_b200 import host
    if not os.path.exists(SPLITFA):
        pytest.skip("oracle/_ref/splitfa not built")
    lengths = [1200, 700, 1500, 100, 2250, 749, 750, 751, 1, 3000]
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, L in enumerate(lengths):
            f.write(">r%d\n" % i)
            s = "T" * L
            for j in range(0, L, 60):
                f.write(s[j:j + 60] + "\n")
    out = subprocess.run([SPLITFA, str(fa), "500"], capture_output=True, text=True, check=True).stdout
    want, name = [], None
    for line in out.splitlines():
        if line.startswith(">"):
            name = line[1:]
            rec, idx = name[1:].split("_")
            want.append([int(rec), int(idx), 0])
        else:
            want[-1][2] += len(line)
    got = host.split_lengths(lengths, 500)
    assert got == [tuple(w) for w in want]


@pytest.mark.parametrize("seed", [0, 1, 7, 12345, 2**31 - 1])
def test_draw_equals_copying_resampler(seed):
    """multiplicities from the srand48(seed) stream == the records `psmc -b --seed` copies (aux.c:8-47)"""
    from psmc_b200 import host
    rng = np.random.default_rng(seed)
    lengths = np.concatenate([np.full(30, 500), rng.integers(500, 750, size=12)]).astype(np.int32)
    mult, mult_copy, view = host.draw_replicate(lengths, seed)
    assert (mult == mult_copy).all()
    assert view[0] == view[3] == mult.sum() and view[1] == view[4] and view[2] == view[5]
    total, drawn = int(lengths.sum()), int((mult * lengths).sum())
    assert abs(drawn - total) <= int(lengths.max())          # as close to the original length as one record allows
    assert (mult == 0).any() and (mult > 1).any()


@pytest.mark.gpu
def test_multiplicity_weighted_estep_matches_oracle_on_expanded_records(oracle):
    from psmc_b200 import EStep, Model, synth
    N = 23
    m = make_model(oracle, N, seed=4)
    seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [900, 40, 1, 1300, 5, 600, 77], 5, miss_frac=0.03, miss_mean=20)
    seqs[4] = seqs[4][:0]            # an empty record keeps its slot in the multiplicity vector
    model = Model.from_dense(m["a0"], m["a"], m["e"])
    with EStep(seqs, N, chunk_len=100) as es:
        base = es.run(model)
        for mult in ([2, 0, 3, 1, 5, 0, 1], [0, 0, 0, 0, 0, 0, 4], [1, 1, 1, 1, 1, 1, 1]):
            es.set_multiplicity(mult)
            got = es.run(model)
            expanded = [s for s, k in zip(seqs, mult) for _ in range(k) if len(s)]
            compare_stats(got, oracle_stats(oracle, m, expanded), 1e-10, N)
            info = es.info()
            assert info["n_seqs_effective"] == sum(k for s, k in zip(seqs, mult) if len(s))
            assert info["active_bins"] == sum(len(s) for s, k in zip(seqs, mult) if k > 0)
        es.set_multiplicity(None)
        compare_stats(es.run(model), base, 1e-12, N)
    # automatic chunk plan (one resident wave) re-planned per replicate
    with EStep(seqs, N) as es:
        es.set_multiplicity([1, 2, 0, 0, 1, 3, 0])
        got = es.run(model)
        expanded = [s for s, k in zip(seqs, [1, 2, 0, 0, 1, 3, 0]) for _ in range(k) if len(s)]
        compare_stats(got, oracle_stats(oracle, m, expanded), 1e-10, N)


@pytest.mark.gpu
def test_multiplicity_rejects_bad_input_and_decode_of_undrawn_records(oracle):
    from psmc_b200 import EStep, Model, Psmc200Error, synth
    N = 23
    m = make_model(oracle, N, seed=4)
    seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [300, 200], 5)
    model = Model.from_dense(m["a0"], m["a"], m["e"])
    with EStep(seqs, N, chunk_len=64) as es:
        with pytest.raises(Psmc200Error):
            es.set_multiplicity([1, -1])
        es.set_multiplicity([0, 2])
        with pytest.raises(Psmc200Error):
            es.decode(model, 0)
        dec = es.decode(model, 1)
        want = oracle.decode(m["a"], m["e"], m["a0"], seqs[1], full=False)
        assert np.max(np.abs(dec["best_p"] - want["best_p"])) < 1e-9


TOL = {"*": (1e-6, 1e-6), "LK": (1e-7, 1e-6), "TR": (5e-5, 2e-6), "MT": (5e-5, 2e-6), "MM": (5e-5, 2e-6),
       "RS": (1e-3, 3e-6), "PA": (1e-3, 3e-6), "RI": (1e-3, 2e-7)}


def _run(args, out):
    r = subprocess.run([PSMC] + args + ["-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return parse(out)


@pytest.mark.gpu
def test_replicates_in_one_process_equal_separate_bootstrap_runs(tmp_path):
    """--replicates R --seed S == cat of R runs `psmc -b --seed S+r` on the split input (README:57-62)"""
    fa = os.path.join(G, "small64.psmcfa.gz")
    common = ["-N2", "-t15", "-r5", "-p", "4+25*2+4+6"]
    # pre-split file through the in-process rule, written back as text by a first run is not needed: --split is applied in both
    batch = _run(common + ["--split=3000", "--replicates", "3", "--seed", "11", fa], str(tmp_path / "batch.psmc"))
    single = []
    for r in range(3):
        single += _run(common + ["--split=3000", "-b", "--seed", str(11 + r), fa], str(tmp_path / ("s%d.psmc" % r)))
    compare_rounds(batch, single, TOL)
    # replicates differ from each other (different draws) and the MM n_seqs line reports the drawn records
    mm = [l for l in batch if l.startswith("MM\tn_seqs")]
    assert len(mm) == 3 and len(set(mm)) > 1
    # one slot / two slots and the order of completion must not change the text
    again = _run(common + ["--split=3000", "--replicates", "3", "--seed", "11", "--slots", "1", fa], str(tmp_path / "b1.psmc"))
    assert again == batch
