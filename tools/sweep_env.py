#!/usr/bin/env python
"""Run bench.py (no CPU baseline) under several environment settings and print one summary line each.
usage: sweep_env.py "A=1 B=2" "A=3" ...   (an empty string = defaults)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
extra = []
if "--scale" in sys.argv:
    i = sys.argv.index("--scale"); extra = ["--scale", sys.argv[i + 1]]; del sys.argv[i:i + 2]
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "12", "--warmup", "5", "--no-cpu", "--no-extras"] + extra, env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        k = d["roofline"]["kernels"]
        fp = d["estep"]["fast_path"]
        print("%-60s it/s %.2f e2e %.2f estep %.2f ms (fwd %.2f [%.2f] bwd %.2f [%.2f]) mstep %.2f fail f/b %d/%d rep %d/%d fb %d" % (
            spec or "(defaults)", d["value"], d["e2e"]["value"], d["roofline"]["estep_ms"], k["forward"]["ms"], k["forward"].get("kernel_alone_ms", 0), k["backward"]["ms"], k["backward"].get("kernel_alone_ms", 0),
            d["estep"]["mstep_ms"], fp["failed_fwd"], fp["failed_bwd"], fp["repaired_fwd"], fp["repaired_bwd"], fp["fallbacks"])
              + " | planned %d ovl %.0f/%.0f slow %d/%d chunks %d x %d rounds %d" % (fp.get("planned", 0), fp.get("avg_overlap_fwd", 0), fp.get("avg_overlap_bwd", 0), fp.get("slow_fwd", 0), fp.get("slow_bwd", 0), fp["n_chunks"], fp["chunk_len"], fp.get("repair_rounds", 0)), flush=True)
    except Exception as e:
        print("%-60s FAILED rc=%d %s %s" % (spec, r.returncode, e, r.stderr[-300:]), flush=True)
