#!/usr/bin/env python
"""tools/make_golden_c2.py -- BASELINE configs[1] golden + a MEASURED reference-vs-reference spread.

  tests/golden/c2.psmcfa.gz   one contig of 500 000 bins drawn from the 64-state PSMC HMM (seeded; bottleneck history)
  tests/golden/c2.psmc        `oracle/_ref/psmc -N25 -t15 -r5 -p 4+25*2+4+6` on it (the UNMODIFIED reference, gcc -O2)
  tests/golden/c2_spread.json how far the reference is from ITSELF when only its floating-point rounding changes:
       * the same sources compiled with FMA contraction (gcc -O2 -mfma -ffp-contract=fast) and with -O3,
       * the same records in a different order (changes only hmm_add_expect's summation order, khmm.c:346-359):
         c2 cut into 5 records by the reference's own splitfa, run as is and reversed.
     Per tag (LK, TR, MT, RS lambda / pi / A columns) the largest relative deviation over all rounds between two such runs
     (tests/psmc_text.py:deviations; one unit in the sixth printed decimal is not counted).
     This is the band inside which "matches the reference" is meaningful; tests/test_cli_gpu.py holds the GPU build to it.

Runs the reference ~5 times for 25 iterations on 500 k bins (about 3-4 minutes each, in parallel).  Dev container only
(needs /root/reference); the variant binaries are built into a temp directory and never enter the repository."""
import gzip
import json
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Ref  # noqa: E402
from psmc_b200 import psmcfa, synth  # noqa: E402
from psmc_text import deviations, parse  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
PAT = "4+25*2+4+6"
ARGS = ["-N25", "-t15", "-r5", "-p", PAT]
SRCS = ["khmm.c", "kmin.c", "cli.c", "core.c", "em.c", "aux.c", "main.c"]


def build_variant(tmp, name, flags):
    out = os.path.join(tmp, "psmc_" + name)
    subprocess.run(["gcc"] + flags + ["-Wno-unused-function", "-Wno-unused-result", "-I" + REF] +
                   [os.path.join(REF, s) for s in SRCS] + ["-o", out, "-lm", "-lz"], check=True, stderr=subprocess.DEVNULL)
    return out


def run(binary, fa, out):
    subprocess.run([binary] + ARGS + ["-o", out, fa], check=True, stderr=subprocess.DEVNULL)
    return parse(out)


def main():
    ref = Ref()
    n, nf, _ = ref.pattern(PAT)
    tm = ref.update_hmm(PAT, np.concatenate([[0.05, 0.0125, 15.0], synth.bottleneck_lambdas(nf)]))
    seq = synth.simulate(tm["a0"], tm["a"], tm["e"], 500000, np.random.default_rng(20260925))
    fa = os.path.join(G, "c2.psmcfa.gz")
    psmcfa.write_psmcfa(fa, [seq], names=["chr22like"])
    with tempfile.TemporaryDirectory() as tmp:
        plain = os.path.join(tmp, "c2.psmcfa")
        psmcfa.write_psmcfa(plain, [seq], names=["chr22like"])
        split = os.path.join(tmp, "c2_split.psmcfa")
        with open(split, "w") as fo:
            subprocess.run([ref.splitfa_bin, plain, "100000"], check=True, stdout=fo)
        recs = open(split).read().split(">")[1:]
        rev = os.path.join(tmp, "c2_split_rev.psmcfa")
        open(rev, "w").write("".join(">" + r for r in reversed(recs)))
        fma = build_variant(tmp, "fma", ["-O2", "-mfma", "-ffp-contract=fast"])
        o3 = build_variant(tmp, "o3", ["-O3"])
        jobs = {"ref": (ref.psmc_bin, plain), "fma": (fma, plain), "o3": (o3, plain),
                "split": (ref.psmc_bin, split), "split_rev": (ref.psmc_bin, rev)}
        with ThreadPoolExecutor(len(jobs)) as ex:
            fut = {k: ex.submit(run, b, f, os.path.join(tmp, k + ".psmc")) for k, (b, f) in jobs.items()}
            res = {k: f.result() for k, f in fut.items()}
        open(os.path.join(G, "c2.psmc"), "w").write("\n".join(res["ref"]) + "\n")
        spread = {"what": "largest relative deviation per tag over all 25 rounds between two runs of the reference that differ only in rounding",
                  "input": "tests/golden/c2.psmcfa.gz (500000 bins, 64 states), psmc " + " ".join(ARGS),
                  "O2_vs_fma_contraction": deviations(res["fma"], res["ref"]),
                  "O2_vs_O3": deviations(res["o3"], res["ref"]),
                  "five_records_vs_reversed_order": deviations(res["split_rev"], res["split"])}
        keys = sorted(set().union(*[set(v) for v in spread.values() if isinstance(v, dict)]))
        spread["max_over_pairs"] = {k: max(v.get(k, 0.0) for v in spread.values() if isinstance(v, dict)) for k in keys}
        json.dump(spread, open(os.path.join(G, "c2_spread.json"), "w"), indent=1)
        print(json.dumps(spread["max_over_pairs"], indent=1))


if __name__ == "__main__":
    main()
