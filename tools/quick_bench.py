#!/usr/bin/env python
"""bench.py (no CPU baseline) with the given steps/warmup under several environment settings; one summary line each.
usage: quick_bench.py STEPS WARMUP "A=1 B=2" "" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps, warm = sys.argv[1], sys.argv[2]
for spec in sys.argv[3:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", warm, "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        k = d["roofline"]["kernels"]; fp = d["estep"]["fast_path"]
        print("%-44s it/s %.2f e2e %.2f estep %.2f ms (fwd %.2f [%.2f] bwd %.2f [%.2f]) mstep %.2f fail f/b %d/%d" % (
            spec or "(defaults)", d["value"], d["e2e"]["value"], d["roofline"]["estep_ms"], k["forward"]["ms"], k["forward"].get("kernel_alone_ms", 0),
            k["backward"]["ms"], k["backward"].get("kernel_alone_ms", 0), d["estep"]["mstep_ms"], fp["failed_fwd"], fp["failed_bwd"]), flush=True)
    except Exception as e:
        print("%-44s FAILED rc=%d %s %s" % (spec, r.returncode, e, r.stderr[-300:]), flush=True)
