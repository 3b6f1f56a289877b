#!/usr/bin/env python
"""Wall time of R bootstrap replicates (BASELINE config 4) through the drop-in binary, all in ONE process:
`psmc --split --replicates R -N<iters> --gpus G` on the synthetic 22-contig genome (splitfa rule in memory, segments
resident once per GPU, a replicate = multiplicity vector, replicates dealt to the GPUs; host/bootstrap.c).

usage: bootstrap_bench.py [--scale S] [--replicates R] [--iters N] [--gpus G[,G2,...]] [--slots K]
Prints one JSON line per GPU count.  The reference recipe (README:57-62) is R separate single-thread processes on the
split file; its cost per replicate is the cost of a plain EM run on the same number of bins (see bench.py cpu_baseline)."""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from psmc_b200 import psmcfa

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--replicates", type=int, default=100)
ap.add_argument("--iters", type=int, default=25)
ap.add_argument("--gpus", default="1")
ap.add_argument("--slots", type=int, default=2)
ap.add_argument("--split", type=int, default=500000)
ap.add_argument("--batch", type=int, default=0, help="replicates per launch sequence (0 = what fits, 1 = one at a time)")
ap.add_argument("--batch-slots", type=int, default=1)
a = ap.parse_args()
seqs = bench.make_genome(a.scale)
fa = "/tmp/genome_boot.psmcfa"
t = time.time(); psmcfa.write_psmcfa(fa, seqs); t_write = time.time() - t
bins = sum(len(s) for s in seqs)
for g in [int(x) for x in a.gpus.split(",")]:
    cmd = [os.path.join(ROOT, "host", "psmc"), "-N%d" % a.iters, "-t15", "-r5", "-p", bench.PATTERN, "--split=%d" % a.split,
           "--replicates", str(a.replicates), "--seed", "1", "--gpus", str(g), "--slots", str(a.slots), "--batch", str(a.batch), "--batch-slots", str(a.batch_slots), "--verbose", "-o", "/tmp/boot.psmc", fa]
    t = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.time() - t
    tail = [l for l in r.stderr.splitlines() if "bootstrap:" in l]
    inner = float(tail[-1].split(":")[-1].split()[0]) if tail else None
    n_rd = sum(1 for l in open("/tmp/boot.psmc") if l.startswith("RD\t%d" % a.iters)) if r.returncode == 0 else 0
    sys.stderr.write("\n".join(r.stderr.splitlines()[-12:]) + "\n")
    print(json.dumps({"metric": "%d-bootstrap wall-time" % a.replicates, "value": dt, "unit": "s", "higher_is_better": False, "n_gpus": g,
                      "replicates": a.replicates, "em_iterations": a.iters, "bins": bins, "split": a.split, "slots_per_gpu": a.slots, "batch": a.batch, "batch_slots": a.batch_slots,
                      "em_phase_s": inner, "replicates_completed": n_rd, "rc": r.returncode,
                      "replicate_iterations_per_s": a.replicates * a.iters / inner if inner else None,
                      "err": r.stderr.strip().splitlines()[-1] if r.returncode else None}), flush=True)
