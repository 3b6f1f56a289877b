#!/usr/bin/env python
"""Throughput of the decode path (BASELINE config 5) through the drop-in binary: psmc -N0 -i params -d on the synthetic genome."""
import os, subprocess, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from psmc_b200 import host, psmcfa, synth
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
seqs = bench.make_genome(scale)
fa = "/tmp/genome.psmcfa"
t = time.time(); psmcfa.write_psmcfa(fa, seqs); print("wrote %s in %.1f s" % (fa, time.time() - t))
n, nf, _ = host.parse_pattern(bench.PATTERN)
par = "/tmp/params.txt"
open(par, "w").write(bench.PATTERN + " " + " ".join("%.9f" % x for x in np.concatenate([[bench.TRUE_THETA, bench.TRUE_RHO, bench.MAX_T], synth.bottleneck_lambdas(nf)])) + "\n")
bins = sum(len(s) for s in seqs)
for flags, tag in ((["-N0", "-i", par], "read+round0"), (["-N0", "-i", par, "-d"], "decode -d"), (["-N1", "-i", par], "1 EM iteration")):
    t = time.time()
    r = subprocess.run([os.path.join(ROOT, "host", "psmc")] + flags + ["-o", "/tmp/out.psmc", "--verbose", fa], capture_output=True, text=True)
    dt = time.time() - t
    print("%-16s rc=%d wall %.2f s  (%.1f Mbin/s)  out %.1f MB  %s" % (tag, r.returncode, dt, bins / dt / 1e6, os.path.getsize("/tmp/out.psmc") / 1e6, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""))
