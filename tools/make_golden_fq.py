#!/usr/bin/env python
"""Fixtures for tests/test_fq2psmcfa.py: small seeded consensus FASTQ / FASTA inputs and what the UNMODIFIED reference utility
(oracle/_ref/fq2psmcfa, built by oracle/Makefile from /root/reference/utils/fq2psmcfa.c) prints for them under every option
the utility has.  Run here (the reference sources do not travel); the outputs are committed under tests/golden/fq/.

usage: python tools/make_golden_fq.py"""
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "fq")
REF = os.path.join(ROOT, "oracle", "_ref", "fq2psmcfa")
REF_SPLIT = os.path.join(ROOT, "oracle", "_ref", "splitfa")

# option sets: every mask rule, quality / block / good-base thresholds, the pseudo-autosomal mask
G = ["-g", "300"]   # (the default, 10000 good bases, drops most of these small records: kept in "default" only)
CASES = {"default": [], "g300": G, "q20": G + ["-q", "20"], "q0": G + ["-q", "0"], "s37": G + ["-s", "37"], "s1": G + ["-s", "1"], "tv": G + ["-v"],
         "ts": G + ["-n"], "cpg_only": G + ["-c"], "cpg_excl": G + ["-C"], "par": G + ["-x"], "q25_s64_C": ["-q", "25", "-s", "64", "-C", "-g", "100"]}


def consensus(rng, n, het=0.02, miss=0.05, lower=0.03):
    """a diploid consensus: mostly ACGT, IUPAC two-allele codes at `het`, runs of N / lower case, a few 3-allele codes, X and '-'"""
    s = rng.choice(list("ACGT"), size=n, p=[0.29, 0.21, 0.21, 0.29])
    # CpG-rich stretches so that the -n / -c / -C rules fire
    for start in rng.integers(0, max(1, n - 40), size=max(1, n // 300)):
        for j in range(start, min(n - 1, start + 30), 2):
            s[j], s[j + 1] = "C", "G"
    idx = rng.random(n) < het
    s[idx] = rng.choice(list("MRWSYK"), size=int(idx.sum()))
    idx = rng.random(n) < 0.002
    s[idx] = rng.choice(list("VHDBX-"), size=int(idx.sum()))
    pos = 0
    while pos < n:          # runs of missing data
        pos += int(rng.geometric(miss / 40.0))
        run = int(rng.geometric(1 / 40.0))
        s[pos:pos + run] = "N"
        pos += run
    idx = rng.random(n) < lower
    s[idx] = np.char.lower(s[idx])
    return "".join(s)


def write_inputs():
    rng = np.random.default_rng(20261017)
    recs = [("chr1", 30050), ("chr2", 12000), ("tiny", 1234), ("chrX", 9000), ("X", 5000), ("exact", 10000), ("mostlyN", 8000), ("empty", 0)]
    fq = []
    for name, n in recs:
        s = consensus(rng, n, miss=0.9 if name == "mostlyN" else 0.05)
        q = "".join(chr(33 + int(v)) for v in rng.choice([2, 8, 12, 19, 20, 24, 30, 40], size=n, p=[.02, .03, .05, .05, .05, .1, .3, .4]))
        # multi-line FASTQ, 73 columns, with a comment in the header
        fq.append("@%s some comment\n" % name)
        fq += [s[i:i + 73] + "\n" for i in range(0, n, 73)]
        fq.append("+\n")
        fq += [q[i:i + 73] + "\n" for i in range(0, n, 73)]
    with gzip.open(os.path.join(OUT, "cons.fq.gz"), "wt", compresslevel=9) as f:
        f.write("".join(fq))
    # plain FASTA (no qualities), single line per record, and a file whose last record has a truncated quality string
    rng = np.random.default_rng(7)
    with open(os.path.join(OUT, "cons.fa"), "w") as f:
        for name, n in (("a", 15000), ("b", 10100)):
            f.write(">%s\n%s\n" % (name, consensus(rng, n)))
    with open(os.path.join(OUT, "trunc.fq"), "w") as f:
        s = consensus(rng, 12000)
        f.write("@ok\n%s\n+\n%s\n" % (s, "I" * 12000))
        s = consensus(rng, 11000)
        f.write("@cut\n%s\n+\n%s\n" % (s, "I" * 9000))
        f.write("@never_reached\n%s\n+\n%s\n" % (s, "I" * 11000))


def main():
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref first: make -C oracle   (needs /root/reference)")
    os.makedirs(OUT, exist_ok=True)
    write_inputs()
    index = {}
    for inp in ("cons.fq.gz", "cons.fa", "trunc.fq"):
        for case, args in CASES.items():
            if inp != "cons.fq.gz" and case not in ("default", "g300", "s37", "ts"):
                continue
            r = subprocess.run([REF] + args + [os.path.join(OUT, inp)], capture_output=True)
            assert r.returncode == 0, r.stderr
            name = "%s.%s.psmcfa" % (inp.replace(".gz", "").replace(".", "_"), case)
            with gzip.open(os.path.join(OUT, name + ".gz"), "wb", compresslevel=9) as f:
                f.write(r.stdout)
            index[name + ".gz"] = {"input": inp, "args": args, "records": r.stdout.count(b">"), "bytes": len(r.stdout)}
    # utils/splitfa.c on two of the files just written (trunk sizes around the 1.5-trunk remainder rule) and on a FASTQ
    for inp, trunks in (("cons_fq.s1.psmcfa.gz", ["2000", "601", "30050", "20034"]), ("cons_fq.g300.psmcfa.gz", ["50", "1"]), ("cons.fq.gz", ["4000"]), ("cons.fa", [])):
        for t in trunks or [None]:
            r = subprocess.run([REF_SPLIT, os.path.join(OUT, inp)] + ([t] if t else []), capture_output=True)
            assert r.returncode == 0, r.stderr
            name = "split.%s.%s.psmcfa.gz" % (inp.replace(".gz", "").replace(".", "_"), t or "default")
            with gzip.open(os.path.join(OUT, name), "wb", compresslevel=9) as f:
                f.write(r.stdout)
            index[name] = {"tool": "splitfa", "input": inp, "args": [t] if t else [], "records": r.stdout.count(b">"), "bytes": len(r.stdout)}
    json.dump(index, open(os.path.join(OUT, "index.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(index, indent=1))


if __name__ == "__main__":
    main()
