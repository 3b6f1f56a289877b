import os, sys, numpy as np, traceback
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle.pyoracle import Oracle
from helpers import make_model, oracle_stats
from psmc_b200 import EStep, Model, synth
o = Oracle()
N = int(os.environ.get("DBG_N", 64))
m = make_model(o, N, seed=3)
seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [int(x) for x in os.environ.get("DBG_L", "3000,500,77").split(",")], seed=5)
want = oracle_stats(o, m, seqs)
mod = Model.from_dense(m["a0"], m["a"], m["e"])
for warm in (0, 100000):
    try:
        with EStep(seqs, N, chunk_len=int(os.environ.get("DBG_CL", 200))) as es:
            es.set_warm(warm)
            got = es.run(mod)
            inf = es.info()
        errs = {"LL": abs(got["LL"] - want["LL"]) / abs(want["LL"])}
        for k in ("E", "RL", "CL", "RU", "CU", "AD"):
            g = np.asarray(got[k]).ravel(); w = np.asarray(want[k]).ravel()
            errs[k] = float(np.max(np.abs(g - w) / np.maximum(np.abs(w), 1e-9 * np.abs(w).max() + 1e-300)))
        print("warm", warm, "G", os.environ.get("PSMC_B200_G_FWD"), os.environ.get("PSMC_B200_G_BWD"), {k: "%.1e" % v for k, v in errs.items()}, "fb", inf["fallbacks"], "rep", inf["repaired_fwd"], inf["repaired_bwd"])
    except Exception as e:
        print("warm", warm, "EXC", repr(e))
