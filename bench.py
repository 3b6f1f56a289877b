#!/usr/bin/env python
"""bench.py -- EM iterations/s of the PSMC hot path on a synthetic 3 Gbp diploid .psmcfa at 64 states.

A "step" is ONE EM iteration over the whole synthetic genome: the E-step (forward/backward/expected
counts over every contig; CUDA, sm_100a) followed by the host M-step (Hooke-Jeeves on the O(N)
objective), exactly what `psmc -N1` adds per round (em.c:27-78).  Workload = BASELINE.json configs[2]
(22 human-autosome-like contigs, 28.8 M bins of 100 bp, pattern 4+25*2+4+6 -> 64 states, -t15 -r5).

  python bench.py --gpus N --steps K --warmup W          own arm (torchrun for N > 1: contigs sharded over
                                                          ranks, one NCCL all-reduce of the 449-double
                                                          statistics vector per iteration)
  python bench.py --impl reference ...                   the reference's own CPU implementation
                                                          (oracle/_ref) on a bounded sample, extrapolated

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATTERN = "4+25*2+4+6"
MAX_T, TR_RATIO = 15.0, 5.0
TRUE_THETA, TRUE_RHO = 0.05, 0.0125
SEED = 20260925
METRIC = "EM iters/sec on 3Gbp psmcfa (n=64)"


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fp:
            d = json.load(fp)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_genome(scale=1.0, lengths=None):
    """seeded synthetic genome drawn from the PSMC HMM itself under a bottleneck history"""
    from psmc_b200 import host, synth
    n, nf, _ = host.parse_pattern(PATTERN)
    params = np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], synth.bottleneck_lambdas(nf)])
    hm = host.model_from_params(PATTERN, params)
    if lengths is None:
        lengths = [max(1000, int(L * scale)) for L in synth.HUMAN_AUTOSOME_BINS]
    t0 = time.time()
    seqs = synth.simulate_genome(hm["a0"], hm["model"].dense(), hm["e"], lengths, SEED)
    log("synthetic genome: %d contigs, %d bins, generated in %.1f s" % (len(seqs), sum(len(s) for s in seqs), time.time() - t0))
    return seqs


from psmc_b200.sharding import lpt_shards  # noqa: E402


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line).  The sampler is started before the
    warm-up (nvidia-smi needs a moment to come up); only rows whose timestamp falls inside [t0, t1] count, and if the timed
    region was too short to catch one, the rows taken while the GPU was under the same load (warm-up .. end) are used."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, smmax, reasons, pw = [], [], set(), []
            for _, r in rows:
                f = [x.strip() for x in r.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); smmax.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, smmax, reasons, pw
        inside = [r for r in self.rows if t0 is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not inside:
            inside, window = self.rows, "warm-up + timed region (the timed region was shorter than one sampling period)"
        sm, smmax, reasons, pw = digest(inside)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the UNMODIFIED reference (oracle/_ref) when present, else the oracle port
# ------------------------------------------------------------------------------------------------
def _cpu_estep_worker(args):
    kind, seed, nbins = args
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Oracle, Ref
    from psmc_b200 import host, synth
    chk = Ref() if kind == "reference" else Oracle()
    n, nf, _ = host.parse_pattern(PATTERN)
    params = np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], synth.bottleneck_lambdas(nf)])
    hm = host.model_from_params(PATTERN, params)
    a = hm["model"].dense()
    seq = synth.simulate(hm["a0"], a, hm["e"], nbins, np.random.default_rng(seed))
    t0 = time.perf_counter()
    chk.estep(a, hm["e"], hm["a0"], [seq])
    return time.perf_counter() - t0


def cpu_baseline(total_bins, cores, sample_bins=100000, reps=1):
    """E-step of the reference on `cores` processes, each on its own sample contig (the contigs of a genome
    are independent, em.c:36-55), plus the reference M-step single-threaded; extrapolated linearly in bins."""
    import multiprocessing as mp
    from oracle.pyoracle import Ref, Oracle
    from psmc_b200 import host, synth
    kind = "reference" if Ref.available() else "port"
    chk = Ref() if kind == "reference" else Oracle()
    t_est = []
    ctx = mp.get_context("spawn")
    for r in range(reps):
        with ctx.Pool(cores) as pool:
            t0 = time.perf_counter()
            per = pool.map(_cpu_estep_worker, [(kind, 1000 + r * 64 + i, sample_bins) for i in range(cores)])
            wall = time.perf_counter() - t0
        t_est.append(max(per))
    t_sample = min(t_est)                                   # seconds for `cores` x sample_bins bins
    bins_per_s = cores * sample_bins / t_sample
    # M-step of the reference: Hooke-Jeeves on the dense O(N^2) objective (em.c:15-25,65), single thread, in C
    n, nf, _ = host.parse_pattern(PATTERN)
    params = np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], np.ones(nf)])
    hm = host.model_from_params(PATTERN, params)
    seq = synth.simulate(hm["a0"], hm["model"].dense(), hm["e"], 20000, np.random.default_rng(5))
    t_m = reference_mstep_seconds(chk, kind, seq)
    t_iter = total_bins / bins_per_s + t_m
    return {"value": 1.0 / t_iter, "unit": "EM iters/s", "cores": cores, "kind": kind,
            "sample": "%d procs x %d bins E-step (%.2f s, %.3g bins/s), M-step %.3f s on 1 core; extrapolated to %d bins"
                      % (cores, sample_bins, t_sample, bins_per_s, t_m, total_bins),
            "estep_bins_per_s": bins_per_s, "mstep_s": t_m, "s_per_iter": t_iter}


def reference_mstep_seconds(chk, kind, seq):
    """time of one M-step of the reference binary: (psmc -N2) - (psmc -N1) - E-step share, on a small input"""
    from psmc_b200 import psmcfa
    import tempfile
    if kind != "reference" or not os.path.exists(chk.psmc_bin):
        return 0.12  # SURVEY.md section 6 probe (0.10-0.15 s); only used when oracle/_ref/psmc is absent
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "s.psmcfa")
        psmcfa.write_psmcfa(fn, [seq])
        def run(nit):
            t0 = time.perf_counter()
            subprocess.run([chk.psmc_bin, "-N%d" % nit, "-t%g" % MAX_T, "-r%g" % TR_RATIO, "-p", PATTERN, "-o", os.path.join(td, "o.psmc"), fn],
                           check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            return time.perf_counter() - t0
        t1 = run(1); t3 = run(3)
        per_iter = (t3 - t1) / 2.0
        # E-step share of that small input, measured through the harness on the same sequence
        from psmc_b200 import host, synth
        n, nf, _ = host.parse_pattern(PATTERN)
        hm = host.model_from_params(PATTERN, np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], np.ones(nf)]))
        t0 = time.perf_counter()
        chk.estep(hm["model"].dense(), hm["e"], hm["a0"], [seq])
        t_e = time.perf_counter() - t0
        return max(per_iter - t_e, 0.02)


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_bins = int(sum(__import__("psmc_b200.synth", fromlist=["x"]).HUMAN_AUTOSOME_BINS) * args.scale)
    cores = min(os.cpu_count() or 1, 64)
    vals = []
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(total_bins, cores, sample_bins=args.cpu_sample_bins)
        if i >= args.warmup:
            vals.append(cb)
        if time.perf_counter() - t_all > 240:
            break
    best = max(vals, key=lambda c: c["value"]) if vals else cb
    v = statistics.mean(c["value"] for c in vals) if vals else cb["value"]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "EM iters/s", "n_gpus": args.gpus, "steps": len(vals) or 1,
           "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(total_bins, args),
           "cpu_baseline": dict(best, value=v),
           "e2e": {"value": v, "unit": "EM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(total_bins, args):
    return {"workload": "configs[2]: 22-contig synthetic diploid genome, %d bins of 100 bp, pattern %s (64 states), -t%g -r%g; "
                        "one step = E-step over all contigs + host M-step" % (total_bins, PATTERN, MAX_T, TR_RATIO),
            "bins": total_bins, "states": 64, "scale": args.scale,
            "l2": "inputs larger than L2 (forward spill %.1f GB per step)" % (total_bins * 520 / 1e9),
            "parallelism": "contigs sharded over %d GPU(s), LPT; NCCL all-reduce of 449 doubles per step" % args.gpus}


def run_own(args):
    import torch
    import torch.distributed as dist
    import ctypes
    import psmc_b200
    from psmc_b200 import host
    from psmc_b200._lib import CInfo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the ONE JSON line and nothing else
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    lib = psmc_b200.load_library()

    seqs = make_genome(args.scale)
    lengths = [len(s) for s in seqs]
    total_bins = sum(lengths)
    owner = lpt_shards(lengths, world)
    mine = [s for s, o in zip(seqs, owner) if o == rank]
    n_seqs_total = len(seqs)
    log("rank %d/%d: %d contigs, %d bins" % (rank, world, len(mine), sum(len(s) for s in mine)))
    sum_L = sum(int((s < 2).sum()) for s in seqs); sum_n = sum(int((s == 1).sum()) for s in seqs)
    theta0 = -np.log(1.0 - sum_n / sum_L)                    # core.c:39 on the WHOLE genome
    n, nf, _ = host.parse_pattern(PATTERN)
    p0 = np.concatenate([[theta0, theta0 / TR_RATIO, MAX_T], np.ones(nf)])
    em = host.EMSession(PATTERN, mine, max_t=MAX_T, tr_ratio=TR_RATIO, init_params=p0, devices=(local,), chunk_len=args.chunk)
    ctx = em.ctx(0)
    slen = lib.psmc_b200_stats_len(ctx)

    class _Dev:  # zero-copy view of the library's device statistics vector for torch.distributed
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
    stats_t = torch.as_tensor(_Dev(lib.psmc_b200_device_stats(ctx), slen), device="cuda:%d" % local)

    kern_ms = []   # per step: library's CUDA-event times of its kernels [K1..K5, total]
    launches = [0]

    def step(upload=False):
        if upload:
            em.upload()                                     # host -> device: 2-bit pack + H2D of every contig
        if world == 1:
            em.estep()                                      # model H2D, kernels, statistics D2H (host buffers in/out)
        else:
            em.launch()
            lib.psmc_b200_wait(ctx)                         # kernels of this rank done (stream sync)
            dist.all_reduce(stats_t)                        # the one collective per EM iteration
            raw = stats_t.cpu().numpy()
            em.set_raw(raw, n_seqs_total)
        ci = CInfo()
        lib.psmc_b200_get_info(ctx, ctypes.byref(ci))
        kern_ms.append(list(ci.ms)[:8]); launches[0] += ci.launches
        em.mstep()                                          # replicated on every rank on identical inputs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, upload):
        barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            step(upload)
        barrier()  # every step already ends with a stream synchronise (statistics D2H); this closes the region on all ranks
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    kern_ms.clear(); launches[0] = 0
    t_clk0 = time.time()
    dt = timed(args.steps, upload=False)
    clocks = sampler.stop(t_clk0, time.time()) if sampler else None
    k_resident = [list(x) for x in kern_ms]
    n_launch = launches[0]
    st = em.state()
    # end to end: every step re-sends all contigs from host memory (pack + H2D), model H2D, statistics D2H
    kern_ms.clear()
    dt_e2e = timed(args.steps, upload=True)
    ci = CInfo()
    lib.psmc_b200_get_info(ctx, ctypes.byref(ci))
    inf = {"n_chunks": ci.n_chunks, "chunk_len": ci.chunk_len, "warm_len": ci.warm_len, "fallbacks": ci.fallbacks,
           "repaired_fwd": ci.repaired_fwd, "repaired_bwd": ci.repaired_bwd, "failed_fwd": ci.failed_fwd, "failed_bwd": ci.failed_bwd,
           "fwd_mismatch": ci.fwd_mismatch, "bwd_mismatch": ci.bwd_mismatch}
    obs_bytes = ci.bytes_obs
    if world > 1:
        t = torch.tensor([float(obs_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        obs_bytes_total = int(t.item())
    else:
        obs_bytes_total = obs_bytes

    if rank == 0:
        peak, peak_src = measured_peaks()
        km = np.array(k_resident)                           # steps x 6
        mean_ms = km.mean(axis=0)
        # algorithmic bytes per bin and EM iteration: forward spill write + re-read, scale factors, 2-bit obs twice
        NST = 64
        alg_bytes_per_bin = 16 * NST + 16.5
        my_bins = sum(len(s) for s in mine)
        estep_ms = float(mean_ms[5])
        dom = int(np.argmax(mean_ms[:5])); names = ["transfer", "chain", "forward", "backward", "reduce"]
        per_kernel = {}
        kb = {"forward": (8 * NST + 8 + 0.25), "backward": (8 * NST + 8 + 0.25)}
        for i, nm in enumerate(names):
            per_kernel[nm] = {"ms": float(mean_ms[i])}
            if nm == "forward" and mean_ms[6] > 0:
                per_kernel[nm]["kernel_alone_ms"] = float(mean_ms[6])      # k_forward without its repair rounds
            if nm == "backward" and mean_ms[7] > 0:
                per_kernel[nm]["kernel_alone_ms"] = float(mean_ms[7])
            if nm in kb and mean_ms[i] > 0:
                per_kernel[nm]["alg_GBps"] = kb[nm] * my_bins / (mean_ms[i] * 1e-3) / 1e9
        achieved = alg_bytes_per_bin * my_bins / (estep_ms * 1e-3) / 1e9
        cores = min(os.cpu_count() or 1, 64)
        cb = cpu_baseline(total_bins, cores, sample_bins=args.cpu_sample_bins) if (world == 1 and not args.no_cpu) else None
        value = args.steps / dt
        out = {"metric": METRIC, "value": value, "unit": "EM iters/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(total_bins, args),
               "e2e": {"value": args.steps / dt_e2e, "unit": "EM iters/s", "h2d_bytes_per_step": int(obs_bytes_total + world * 8 * 64 * 8),
                       "d2h_bytes_per_step": int(world * slen * 8), "ms_per_step": dt_e2e / args.steps * 1e3,
                       "what": "psmch_em_iterate from host buffers; every step re-packs and re-sends all contigs (H2D), sends the model, reads the statistics back"},
               "gpu_launches": int(n_launch),
               "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            # dram__bytes_read+write of k_forward + k_backward per bin from the committed ncu --set full capture
                            # (profiles/r01_gen2_ncu_full.txt: 3.687 GB + 3.759 GB over 7 187 491 bins), scaled to this launch
                            "traffic": 1036.0 * my_bins / 1e9, "traffic_unit": "GB per E-step (k_forward + k_backward; ncu, scaled per bin)",
                            "kernel": "whole E-step (all kernels of one iteration on rank 0; dominant: %s)" % names[dom],
                            "algorithmic_bytes_per_bin": alg_bytes_per_bin, "bins_per_launch": my_bins, "peak_source": peak_src,
                            "estep_ms": estep_ms, "kernels": per_kernel},
               "estep": {"bins_per_s": total_bins / (estep_ms * 1e-3) if world == 1 else None, "ms": estep_ms,
                         "mstep_ms": st["t_mstep_ms"], "hj_calls": st["hj_calls"], "chunks": inf["n_chunks"], "chunk_len": inf["chunk_len"], "fast_path": inf},
               "final": {"lk": st["lk"], "theta": float(st["params"][0]), "rho": float(st["params"][1])}}
        if cb:
            out["cpu_baseline"] = cb
        print(json.dumps(out), flush=True)
    em.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale every contig length (1.0 = the 28.8 M-bin workload)")
    ap.add_argument("--chunk", type=int, default=0, help="bins per chunk (0 = auto)")
    ap.add_argument("--cpu-sample-bins", type=int, default=100000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
