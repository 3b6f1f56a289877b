#!/usr/bin/env python
"""Small run of every device path for `compute-sanitizer` (memcheck / racecheck / synccheck): E-step on the fast path with
repairs and the mixing probe, planned plans, staged (bulk-async + mbarrier) and ring backward, batch mode, dense counts,
whole-context decode, one-sequence decode.  Checked against the oracle so that a sanitizer run is also a parity run.
usage: compute-sanitizer --tool memcheck python tools/sanitize_paths.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import compare_stats, make_model, oracle_stats  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402
from psmc_b200 import EStep, Model, synth  # noqa: E402

o = Oracle()
N = 64
m = make_model(o, N, seed=3)
seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [9000, 2500, 1, 700], 5, miss_frac=0.03, miss_mean=20)
seqs[0][3000:5200] = 0
want = oracle_stats(o, m, seqs)
mod = Model.from_dense(m["a0"], m["a"], m["e"])
for tma in ("1", "0"):
    os.environ["PSMC_B200_TMA"] = tma
    with EStep(seqs, N) as es:                 # automatic plan: probe on E-step 0, planned overlaps afterwards
        es.set_warm(400)
        for it in range(4):
            compare_stats(es.run(mod), want, 1e-10, N)
        print("tma", tma, {k: es.info()[k] for k in ("planned", "n_chunks", "failed_fwd", "failed_bwd", "repair_rounds", "fallbacks")})
        es.set_dense(True)
        compare_stats(es.run(mod), want, 1e-10, N)
        A = es.dense_counts()
        assert np.max(np.abs(A - want["A"]) / np.maximum(np.abs(want["A"]), 1e-9 * np.abs(want["A"]).max())) < 1e-10
os.environ["PSMC_B200_TMA"] = "1"
with EStep(seqs, N, chunk_len=300) as es:
    es.set_warm(350)
    mults = np.array([[1, 0, 2, 1], [0, 3, 1, 1]], dtype=np.int32)
    es.set_batch(mults)
    got = es.run_batch([mod, mod])
    for r in range(2):
        compare_stats(got[r], oracle_stats(o, m, [s for s, k in zip(seqs, mults[r]) for _ in range(k)]), 1e-10, N)
    es.set_multiplicity(None)
    dec = es.decode_all(mod, runs=True, bins=True, post=True)
    w = o.decode(m["a"], m["e"], m["a0"], seqs[0], full=True)
    assert np.max(np.abs(dec["seqs"][0]["post"] - w["post"])) < 2e-7
    one = es.decode(mod, 1, full=True, want_s=True)
    w1 = o.decode(m["a"], m["e"], m["a0"], seqs[1], full=True)
    assert np.max(np.abs(one["post"] - w1["post"])) < 1e-10
print("sanitize_paths: all paths ran and match the oracle")
