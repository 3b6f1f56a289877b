"""N > 1 path on CPU: world_size-2 gloo.  Contigs are sharded with the product's LPT rule, every rank produces the
raw statistics vector of ITS shard (here with the oracle standing in for the GPU kernels -- test infrastructure),
the vectors are summed by ONE all-reduce, unpacked by the C ABI (psmc_b200_unpack_stats) and fed to the replicated
host M-step.  Result must equal the single-process run and be identical on both ranks."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_raw(oracle, m, seqs):
    from psmc_b200.sharding import pack_raw
    if not seqs:
        return np.zeros(7 * m["N"] + 1)
    r = oracle.estep(m["a"], m["e"], m["a0"], seqs)
    S = oracle.struct_stats(r["A"])
    return pack_raw(r["LL"], r["E"], S["RL"], S["CL"], S["RU"], S["CU"], S["AD"])


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from psmc_b200 import host, synth
    from psmc_b200.estep import _StatsBuf
    from psmc_b200._lib import load_library
    from psmc_b200.sharding import lpt_shards, all_reduce_raw
    from helpers import make_model
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = Oracle()
    m = make_model(o, 23, seed=9)
    seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [3000, 800, 2500, 1200, 50], seed=10)
    owner = lpt_shards([len(s) for s in seqs], world)
    mine = [s for s, w in zip(seqs, owner) if w == rank]
    raw = torch.from_numpy(_shard_raw(o, m, mine))
    all_reduce_raw(raw)
    lib = load_library()
    buf = _StatsBuf(m["N"])
    import ctypes as C
    arr = np.ascontiguousarray(raw.numpy())
    assert lib.psmc_b200_unpack_stats(m["N"], arr.ctypes.data_as(C.POINTER(C.c_double)), len(seqs), C.byref(buf.c)) == 0
    st = buf.result()
    res = host.mstep(m["pattern"], m["params"], st["E"], marg=st)
    # the exchange bench.py uses from round 2 on: the M-step's result of rank 0 is broadcast, the other ranks rebuild their
    # model from it (here: every rank also ran the search, so the broadcast value can be checked against the local one)
    from psmc_b200.sharding import broadcast_params
    bp = torch.from_numpy(res["params"].copy() if rank == 0 else np.zeros_like(res["params"]))
    broadcast_params(bp, 0)
    mod_b = host.model_from_params(m["pattern"], bp.numpy())["model"]
    mod_l = host.model_from_params(m["pattern"], res["params"])["model"]
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), LL=st["LL"], E=st["E"], RL=st["RL"], params=res["params"], Q1=res["Q1"],
             owner=np.array(owner), bparams=bp.numpy(), same_model=np.array([np.array_equal(mod_b.U, mod_l.U) and np.array_equal(mod_b.D, mod_l.D)]))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process(oracle, tmp_path):
    import torch.multiprocessing as mp
    from psmc_b200 import host, synth
    from helpers import make_model
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("LL", "E", "RL", "params", "Q1"):
        assert np.array_equal(r0[k], r1[k]), k                     # all-reduce result and M-step identical on both ranks
    assert set(r0["owner"].tolist()) == {0, 1}
    assert np.array_equal(r1["bparams"], r0["params"]) and bool(r1["same_model"][0])   # rank 1 received rank 0's parameters
    m = make_model(oracle, 23, seed=9)
    seqs = synth.simulate_genome(m["a0"], m["a"], m["e"], [3000, 800, 2500, 1200, 50], seed=10)
    one = oracle.estep(m["a"], m["e"], m["a0"], seqs)
    S = oracle.struct_stats(one["A"])
    assert abs(r0["LL"] - one["LL"]) <= 1e-12 * abs(one["LL"])
    assert np.allclose(r0["E"], one["E"], rtol=1e-12) and np.allclose(r0["RL"][1:], S["RL"][1:], rtol=1e-12)
    ref = host.mstep(m["pattern"], m["params"], one["E"], marg=S)
    assert abs(r0["Q1"] - ref["Q1"]) <= 1e-9 * abs(ref["Q1"])


def test_lpt_shards_balance():
    from psmc_b200.sharding import lpt_shards
    from psmc_b200.synth import HUMAN_AUTOSOME_BINS
    for n in (1, 2, 4, 8):
        owner = lpt_shards(HUMAN_AUTOSOME_BINS, n)
        load = [sum(L for L, o in zip(HUMAN_AUTOSOME_BINS, owner) if o == r) for r in range(n)]
        assert max(load) <= 1.12 * sum(load) / n
