#!/usr/bin/env python
"""tools/make_golden.py -- generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref, built from /root/reference by oracle/Makefile) on seeded synthetic inputs.

The reference owns no tests or fixtures, so these are made here and committed:
  c1.psmcfa.gz / c1.psmc         BASELINE config 1: 10 000 bins, -N5 -t5 -r1 -p 4+5*3+4 (23 states), reference output
  c1_params.txt                  the PA line of c1.psmc's last round (input of -i)
  c1_decode.psmc                 -N0 -i c1_params.txt -d (TC + DC lines)
  c1_prob.psmc                   -N0 -i c1_params.txt -s (PR line)
  c1_fulldecode.psmc.gz          -N0 -i c1_params.txt -D (DF lines)
  small64.psmcfa.gz / .psmc      3 contigs (30k bins), -N4 -t15 -r5 -p 4+25*2+4+6 (64 states)
  estep_*.npz                    dense E-step results of the reference (LL, A, E) for fixed models, through ref_harness.c
Run in the dev container only (needs /root/reference):  python tools/make_golden.py
"""
import gzip
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402
from psmc_b200 import psmcfa, synth  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def run_ref(ref, args, out):
    subprocess.run([ref.psmc_bin] + args + ["-o", out], check=True, stderr=subprocess.DEVNULL)


def main():
    os.makedirs(G, exist_ok=True)
    ref = Ref()
    # ---- C1
    pat = "4+5*3+4"
    n, nf, _ = ref.pattern(pat)
    tm = ref.update_hmm(pat, np.concatenate([[0.04, 0.01, 5.0], synth.bottleneck_lambdas(nf)]))
    seqs = synth.simulate_genome(tm["a0"], tm["a"], tm["e"], [10000], seed=101)
    psmcfa.write_psmcfa(os.path.join(G, "c1.psmcfa.gz"), seqs, names=["chrA"])
    fa = os.path.join(G, "c1.psmcfa.gz")
    run_ref(ref, ["-N5", "-t5", "-r1", "-p", pat, fa], os.path.join(G, "c1.psmc"))
    # decoding with FIXED parameters (-N0 -i): the PA line of the last round, minus its tag (aux.c:84-113)
    pa = [l for l in open(os.path.join(G, "c1.psmc")) if l.startswith("PA\t")][-1][3:]
    par = os.path.join(G, "c1_params.txt")
    open(par, "w").write(pa)
    run_ref(ref, ["-N0", "-i", par, "-d", fa], os.path.join(G, "c1_decode.psmc"))
    run_ref(ref, ["-N0", "-i", par, "-s", fa], os.path.join(G, "c1_prob.psmc"))
    tmp = os.path.join(G, "c1_fulldecode.psmc")
    run_ref(ref, ["-N0", "-i", par, "-D", fa], tmp)
    with open(tmp, "rb") as fi, gzip.open(tmp + ".gz", "wb") as fo:
        fo.write(fi.read())
    os.remove(tmp)
    # ---- 64 states, 3 ragged contigs
    pat = "4+25*2+4+6"
    n, nf, _ = ref.pattern(pat)
    tm = ref.update_hmm(pat, np.concatenate([[0.05, 0.0125, 15.0], synth.bottleneck_lambdas(nf)]))
    seqs = synth.simulate_genome(tm["a0"], tm["a"], tm["e"], [17000, 9000, 4000], seed=102)
    fa = os.path.join(G, "small64.psmcfa.gz")
    psmcfa.write_psmcfa(fa, seqs, names=["c1", "c2", "c3"])
    run_ref(ref, ["-N4", "-t15", "-r5", "-p", pat, fa], os.path.join(G, "small64.psmc"))
    # ---- dense E-step vectors through the harness (hmm_forward/backward/expect of the reference itself)
    for tag, pat, lens, seed in (("23", "4+5*3+4", [3000, 1, 2, 517], 7), ("64", "4+25*2+4+6", [2500, 700], 8)):
        n, nf, _ = ref.pattern(pat)
        params = np.concatenate([[0.05, 0.0125, 15.0], synth.bottleneck_lambdas(nf) * 1.1])
        m = ref.update_hmm(pat, params)
        seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], lens, seed=seed, miss_frac=0.03, miss_mean=20)
        r = ref.estep(m["a"], m["e"], m["a0"], seqs)
        np.savez_compressed(os.path.join(G, "estep_%s.npz" % tag), pattern=pat, params=params, a=m["a"], e=m["e"], a0=m["a0"],
                            sigma=m["sigma"], t=m["t"], C_pi=m["C_pi"], C_sigma=m["C_sigma"],
                            seqs=np.concatenate(seqs), lens=np.array(lens), LL=r["LL"], A=r["A"], E=r["E"], Q0=r["Q0"])
    print("golden fixtures written to", G)
    for f in sorted(os.listdir(G)):
        print("  %-28s %8d bytes" % (f, os.path.getsize(os.path.join(G, f))))


if __name__ == "__main__":
    main()
