#!/usr/bin/env python
"""bench.py -- EM iterations/s of the PSMC hot path on a synthetic 3 Gbp diploid .psmcfa at 64 states.

A "step" is ONE EM iteration over the whole synthetic genome: the E-step (forward/backward/expected
counts over every contig; CUDA, sm_100a) followed by the host M-step (Hooke-Jeeves on the O(N)
objective), exactly what `psmc -N1` adds per round (em.c:27-78).  Workload = BASELINE.json configs[2]
(22 human-autosome-like contigs, 28.8 M bins of 100 bp, pattern 4+25*2+4+6 -> 64 states, -t15 -r5).

  python bench.py --gpus N --steps K --warmup W          own arm (torchrun for N > 1: contigs sharded over
                                                          ranks, one NCCL all-reduce of the 449-double
                                                          statistics vector per iteration, M-step on rank 0,
                                                          NCCL broadcast of the parameters)
  python bench.py --impl reference ...                   the reference's own CPU implementation (oracle/_ref):
                                                          every step times its E-step on a bounded sample on
                                                          all host cores; value = the whole-workload rate that
                                                          follows; plus a measured (not extrapolated) 1-core run
                                                          of the unmodified binary on the configs[1] contig
  python bench.py --workload bootstrap|decode            BASELINE's second metric (100-bootstrap wall time,
                                                          configs[3]) / configs[4] (decode throughput) alone;
                                                          the default run appends both as "bootstrap" / "decode"
                                                          objects to its line (--no-extras skips them)

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


from psmc_b200.sharding import broadcast_params, lpt_shards  # noqa: E402


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
# CPU baseline / reference arm: the UNMODIFIED reference (oracle/_ref) when present, else the oracle port.
# Nothing in here touches the product libraries (no psmc_b200.host, no libpsmc_b200.so): the model comes from the
# reference's own psmc_update_hmm (core.c:61-133 through oracle/ref_harness.c), the data from numpy.
# ------------------------------------------------------------------------------------------------
def _checker():
    from oracle.pyoracle import Ref, Oracle
    return (Ref(), "reference") if Ref.available() else (Oracle(), "port")


def _true_model(chk):
    from psmc_b200 import synth
    n, nf, _ = chk.pattern(PATTERN)
    return chk.update_hmm(PATTERN, np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], synth.bottleneck_lambdas(nf)]))


def _cpu_estep_worker(args):
    seed, nbins = args
    sys.path.insert(0, ROOT)
    from psmc_b200 import synth
    chk, _ = _checker()
    m = _true_model(chk)
    seq = synth.simulate(m["a0"], m["a"], m["e"], nbins, np.random.default_rng(seed))
    t0 = time.perf_counter()
    chk.estep(m["a"], m["e"], m["a0"], [seq])
    return time.perf_counter() - t0


class CpuArm:
    """E-step of the reference (hmm_forward/backward/lk/expect, khmm.c:145-324, through its own objects) on sample
    contigs: `cores` processes at once, each on its own sample (the contigs of a genome are independent, em.c:36-55)."""

    def __init__(self, cores, sample_bins):
        import multiprocessing as mp
        self.cores, self.sample_bins = cores, sample_bins
        self.pool = mp.get_context("spawn").Pool(cores)
        self.pool.map(_cpu_estep_worker, [(1, 2000)] * cores)          # start the workers, load the library
        self.round = 0

    def step(self):
        """one timed all-core sample E-step; returns (wall seconds, bins/s over all cores)"""
        self.round += 1
        t0 = time.perf_counter()
        per = self.pool.map(_cpu_estep_worker, [(1000 + self.round * 64 + i, self.sample_bins) for i in range(self.cores)])
        wall = time.perf_counter() - t0
        return wall, self.cores * self.sample_bins / max(per)

    def one_core(self):
        """the same E-step on ONE core (the reference is single-threaded: BASELINE.md section 3's headline figure)"""
        t = _cpu_estep_worker((999, self.sample_bins))
        return self.sample_bins / t

    def close(self):
        self.pool.close()
        self.pool.join()


def _ref_binary_runs(chk, seq, n_list, pattern=PATTERN, extra=()):
    """wall seconds of `oracle/_ref/psmc -N n` on seq for every n of n_list (unmodified binary, one core)"""
    from psmc_b200 import psmcfa
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "s.psmcfa")
        psmcfa.write_psmcfa(fn, [seq])
        for n in n_list:
            t0 = time.perf_counter()
            subprocess.run([chk.psmc_bin, "-N%d" % n, "-t%g" % MAX_T, "-r%g" % TR_RATIO, "-p", pattern] + list(extra) +
                           ["-o", os.path.join(td, "o.psmc"), fn], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            out[n] = time.perf_counter() - t0
    return out


def reference_mstep_seconds(chk, kind):
    """one M-step of the reference (Hooke-Jeeves on the dense O(N^2) objective, em.c:15-25,65; independent of the input
    length): per-iteration time of the unmodified binary on a small input minus the E-step share of that input"""
    from psmc_b200 import synth
    if kind != "reference" or not os.path.exists(chk.psmc_bin):
        return 0.12  # SURVEY.md section 6 probe (0.10-0.15 s); only used when oracle/_ref/psmc is absent
    m = _true_model(chk)
    seq = synth.simulate(m["a0"], m["a"], m["e"], 20000, np.random.default_rng(5))
    t = _ref_binary_runs(chk, seq, (1, 3))
    per_iter = (t[3] - t[1]) / 2.0
    flat = chk.update_hmm(PATTERN, np.concatenate([[TRUE_THETA, TRUE_RHO, MAX_T], np.ones(chk.pattern(PATTERN)[1])]))
    t0 = time.perf_counter()
    chk.estep(flat["a"], flat["e"], flat["a0"], [seq])
    return max(per_iter - (time.perf_counter() - t0), 0.02)


def measured_c2(chk, kind, bins=500000, k=2):
    """MEASURED, not extrapolated: (T(-N k) - T(-N 0)) / k of the unmodified binary on the configs[1] contig"""
    from psmc_b200 import synth
    if kind != "reference" or not os.path.exists(chk.psmc_bin):
        return None
    m = _true_model(chk)
    seq = synth.simulate(m["a0"], m["a"], m["e"], bins, np.random.default_rng(SEED))
    t = _ref_binary_runs(chk, seq, (0, k))
    return {"bins": bins, "iterations": k, "s_per_iter": (t[k] - t[0]) / k, "wall_s": {"N0": t[0], "N%d" % k: t[k]}, "cores": 1}


def cpu_baseline(total_bins, cores, sample_bins=100000, with_c2=True, arm=None):
    """the reported CPU baseline of the own arm (rank 0, N = 1): one all-core sample E-step, the 1-core figure, the
    reference M-step, and the measured configs[1] iteration of the unmodified binary"""
    chk, kind = _checker()
    own = arm is None
    arm = arm or CpuArm(cores, sample_bins)
    wall, bps = arm.step()
    bps1 = arm.one_core()
    if own:
        arm.close()
    t_m = reference_mstep_seconds(chk, kind)
    t_iter = total_bins / bps + t_m
    t_iter1 = total_bins / bps1 + t_m
    out = {"value": 1.0 / t_iter, "unit": "EM iters/s", "cores": cores, "kind": kind,
           "sample": "%d procs x %d bins E-step (%.2f s, %.3g bins/s), M-step %.3f s on 1 core; extrapolated linearly in bins to %d bins"
                     % (cores, sample_bins, wall, bps, t_m, total_bins),
           "estep_bins_per_s": bps, "mstep_s": t_m, "s_per_iter": t_iter,
           "one_core": {"value": 1.0 / t_iter1, "unit": "EM iters/s", "estep_bins_per_s": bps1, "s_per_iter": t_iter1,
                        "what": "the reference is single-threaded: this is what one `psmc` process does (BASELINE.md section 3)"}}
    if with_c2:
        c2 = measured_c2(chk, kind)
        if c2:
            c2["extrapolated_s_per_iter"] = c2["bins"] / bps1 + t_m      # what the 1-core sample predicts for the same contig
            c2["measured_over_extrapolated"] = c2["s_per_iter"] / c2["extrapolated_s_per_iter"]
            out["measured_configs1"] = c2
    return out


def bootstrap_cpu_baseline(cb, bins, replicates, iters):
    """README:57-62 recipe: R independent single-thread `psmc -b` runs over the split file, `xargs -P cores` of them at
    a time.  Derived from the measured 1-core E-step rate and M-step time (a replicate has as many bins as the genome)."""
    per_rep = iters * (bins / cb["one_core"]["estep_bins_per_s"] + cb["mstep_s"])
    return {"one_core_s": replicates * per_rep, "all_cores_s": replicates * per_rep / max(1, min(cb["cores"], replicates)), "cores": cb["cores"],
            "kind": cb["kind"], "what": "derived: %d replicates x %d iterations x (bins / 1-core E-step rate + M-step); all cores = xargs -P %d" % (replicates, iters, cb["cores"])}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from psmc_b200 import synth
    total_bins = int(sum(synth.HUMAN_AUTOSOME_BINS) * args.scale)
    cores = min(os.cpu_count() or 1, 64)
    chk, kind = _checker()
    t_all = time.perf_counter()
    arm = CpuArm(cores, args.cpu_sample_bins)
    t_m = reference_mstep_seconds(chk, kind)
    walls, rates = [], []
    for i in range(args.warmup + args.steps):
        wall, bps = arm.step()                      # one step = one all-core sample E-step (+ the M-step time measured above)
        if i >= args.warmup:
            walls.append(wall); rates.append(bps)
        if time.perf_counter() - t_all > 150:
            break
    bps1 = arm.one_core()
    arm.close()
    if not rates:
        rates, walls = [bps], [wall]
    bps = statistics.mean(rates)
    t_iter = total_bins / bps + t_m
    v = 1.0 / t_iter
    cb = {"value": v, "unit": "EM iters/s", "cores": cores, "kind": kind,
          "sample": "%d steps; each: %d procs x %d bins E-step of the reference (mean %.2f s, %.3g bins/s over all cores); M-step %.3f s (binary, 1 core); "
                    "value = 1 / (bins / rate + M-step) for the %d-bin workload" % (len(rates), cores, args.cpu_sample_bins, statistics.mean(walls), bps, t_m, total_bins),
          "estep_bins_per_s": bps, "mstep_s": t_m, "s_per_iter": t_iter,
          "one_core": {"value": 1.0 / (total_bins / bps1 + t_m), "unit": "EM iters/s", "estep_bins_per_s": bps1, "s_per_iter": total_bins / bps1 + t_m}}
    c2 = measured_c2(chk, kind) if not args.no_extras else None
    if c2:
        c2["extrapolated_s_per_iter"] = c2["bins"] / bps1 + t_m
        c2["measured_over_extrapolated"] = c2["s_per_iter"] / c2["extrapolated_s_per_iter"]
        cb["measured_configs1"] = c2
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "EM iters/s", "n_gpus": args.gpus, "steps": len(rates),
           "warmup": args.warmup, "ms_per_step": statistics.mean(walls) * 1e3,
           "ms_per_step_what": "wall time of one TIMED step = the bounded sample E-step on all cores; the whole-workload iteration it implies is extrapolated_ms_per_iteration",
           "extrapolated_ms_per_iteration": t_iter * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(total_bins, args),
           "cpu_baseline": cb,
           "bootstrap": bootstrap_cpu_baseline(cb, total_bins, args.boot_replicates, args.boot_iters),
           "e2e": {"value": v, "unit": "EM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def workload_config(total_bins, args):
    return {"workload": "configs[2]: 22-contig synthetic diploid genome, %d bins of 100 bp, pattern %s (64 states), -t%g -r%g; "
                        "one step = E-step over all contigs + host M-step" % (total_bins, PATTERN, MAX_T, TR_RATIO),
            "bins": total_bins, "states": 64, "scale": args.scale,
            "l2": "inputs larger than L2 (forward spill %.1f GB per step)" % (total_bins * 520 / 1e9),
            "parallelism": "contigs sharded over %d GPU(s), LPT; NCCL all-reduce of 449 doubles per step, M-step on rank 0, NCCL broadcast of the parameters" % args.gpus}


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
    # stdout carries the ONE JSON line and nothing else: NCCL prints its version banner to file descriptor 1 on some hosts
    # whatever NCCL_DEBUG_FILE says, so descriptor 1 points at stderr until the line is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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

    par_t = torch.zeros(em.n_params, dtype=torch.float64, device="cuda:%d" % local)
    kern_ms = []   # per step: library's CUDA-event times of its kernels [K1..K5, total]
    launches = [0]

    pending = [None]   # e2e: the upload of the NEXT step's contigs, started while this step's M-step runs on the host

    def step(upload=False, prefetch=False):
        if upload:
            if pending[0] is not None:                      # this step's inputs were sent while the previous M-step ran
                pending[0].join(); pending[0] = None
                if up_err:
                    raise up_err[0]
            else:
                em.upload()                                 # host -> device: 2-bit pack + H2D of every contig
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
        if upload and prefetch:
            # the E-step's kernels are done (statistics read back): the device buffers are free, and the M-step needs no GPU --
            # copy the next step's inputs under it (pack + H2D, same bytes every step, still inside the timed region)
            pending[0] = threading.Thread(target=_upload_bg)
            pending[0].start()
        if world == 1:
            em.mstep()
        else:
            # the M-step runs on rank 0 only (with its helper threads); the other ranks receive the parameters: 8 replicated
            # searches on one host fought for its cores (round 1: 4.6 ms per M-step at 8 ranks against 3.2 ms alone)
            if rank == 0:
                em.mstep()
                par_t.copy_(torch.from_numpy(em.state()["params"]))
            broadcast_params(par_t, 0)
            if rank != 0:
                em.set_params(par_t.cpu().numpy())

    up_err = []

    def _upload_bg():
        try:
            em.upload()
        except Exception as e:   # surfaces at the join
            up_err.append(e)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, upload):
        barrier()
        t0 = time.perf_counter()
        for i in range(k):
            step(upload, prefetch=(i + 1 < k))
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
           "fwd_mismatch": ci.fwd_mismatch, "bwd_mismatch": ci.bwd_mismatch,
           "planned": ci.planned, "probe_plans": ci.probe_plans, "avg_overlap_fwd": ci.avg_overlap_fwd, "avg_overlap_bwd": ci.avg_overlap_bwd,
           "slow_fwd": ci.slow_fwd, "slow_bwd": ci.slow_bwd, "repair_rounds": ci.repair_rounds, "warm_redos": ci.warm_redos,
           "n_chunks_bwd": ci.n_chunks_bwd, "chunk_len_bwd": ci.chunk_len_bwd}
    obs_bytes = ci.bytes_obs
    if world > 1:
        t = torch.tensor([float(obs_bytes)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        obs_bytes_total = int(t.item())
    else:
        obs_bytes_total = obs_bytes

    out, cb = None, None
    if rank == 0:
        peak, peak_src = measured_peaks()
        km = np.array(k_resident)                           # steps x 6
        mean_ms = km.mean(axis=0)
        # algorithmic bytes per bin and EM iteration: forward spill write + re-read, scale factors, 2-bit obs twice
        NST = 64
        alg_bytes_per_bin = 16 * NST + 16.5
        FP64_LANE_INSTR_PER_BIN = 122 * 32 // 4 + 130 * 32 // 2     # 976 forward + 2080 backward
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
        cb = cpu_baseline(total_bins, cores, sample_bins=args.cpu_sample_bins, with_c2=not args.no_extras) if (world == 1 and not args.no_cpu) else None
        value = args.steps / dt
        out = {"metric": METRIC, "value": value, "unit": "EM iters/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": workload_config(total_bins, args),
               "e2e": {"value": args.steps / dt_e2e, "unit": "EM iters/s", "h2d_bytes_per_step": int(obs_bytes_total + world * 8 * 64 * 8),
                       "d2h_bytes_per_step": int(world * slen * 8), "ms_per_step": dt_e2e / args.steps * 1e3,
                       "what": "EM iteration from host buffers; every step re-packs and re-sends all contigs (H2D), sends the model, reads the statistics back; "
                               "the pack + H2D of step i+1 runs on a host thread while the M-step of step i runs (no GPU work in flight then)"},
               "gpu_launches": int(n_launch),
               "clocks": clocks,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            # dram__bytes_read+write of k_forward + k_backward per bin from the committed ncu --set full capture
                            # (profiles/r02_ncu_full.txt: k_forward 3.680 GB written + k_backward_staged 3.755 GB read over 7 187 491 bins), scaled to this launch
                            "traffic": 1036.0 * my_bins / 1e9, "traffic_unit": "GB per E-step (k_forward + k_backward; ncu, scaled per bin)",
                            "kernel": "whole E-step (all kernels of one iteration on rank 0; dominant: %s)" % names[dom],
                            "algorithmic_bytes_per_bin": alg_bytes_per_bin, "bins_per_launch": my_bins, "peak_source": peak_src,
                            "estep_ms": estep_ms, "kernels": per_kernel,
                            # the pipe that actually limits these kernels: FP64 instructions per lane and bin counted in the
                            # SASS of the two hot loops (k_forward<8,8,2> stored bin: 122 warp instructions per 4 chunks; k_backward_staged<4,16>:
                            # 130 of the 216 per 2 chunks; overlap and repair work NOT counted), against the DFMA issue rate measured on B200
                            # (profiles/r01_ubench_b200.txt: 34 TFLOP/s = 17e12 lane instructions/s)
                            "fp64": {"lane_instr_per_bin": FP64_LANE_INSTR_PER_BIN, "achieved_Tinstr_s": FP64_LANE_INSTR_PER_BIN * my_bins / (estep_ms * 1e-3) / 1e12,
                                     "peak_Tinstr_s": 17.0, "frac": FP64_LANE_INSTR_PER_BIN * my_bins / (estep_ms * 1e-3) / 17.0e12,
                                     "pipe_active_pct_ncu": {"k_forward": 51.0, "k_backward_staged": 56.5, "source": "profiles/r02_ncu_full.txt"}}},
               "estep": {"bins_per_s": total_bins / (estep_ms * 1e-3) if world == 1 else None, "ms": estep_ms,
                         "mstep_ms": st["t_mstep_ms"], "hj_calls": st["hj_calls"], "chunks": inf["n_chunks"], "chunk_len": inf["chunk_len"], "fast_path": inf},
               "final": {"lk": st["lk"], "theta": float(st["params"][0]), "rho": float(st["params"][1])}}
        if cb:
            out["cpu_baseline"] = cb
    em.close()
    del stats_t
    torch.cuda.empty_cache()
    # ---- secondary metrics through the drop-in binary (one process drives all N GPUs; the other ranks stay idle on the CPU)
    store = dist.distributed_c10d._get_default_store() if world > 1 else None
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    if rank == 0:
        if not args.no_extras:
            fa = write_genome_file(seqs)
            try:
                out["bootstrap"] = run_bootstrap(args, fa, total_bins, world, cb)
            except Exception as e:      # the headline line must survive a failing extra
                out["bootstrap"] = {"error": str(e)[:300]}
            try:
                out["decode"] = run_decode(args, fa, total_bins, world)
            except Exception as e:
                out["decode"] = {"error": str(e)[:300]}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
        if store is not None:
            store.set("bench_extras_done", "1")
    elif store is not None:
        store.wait(["bench_extras_done"])     # blocks on the CPU: no NCCL kernel spins on this rank's GPU meanwhile
    if world > 1:
        dist.destroy_process_group()


GENOME_FA = "/tmp/psmc_b200_bench_genome.psmcfa"


def write_genome_file(seqs):
    from psmc_b200 import psmcfa
    psmcfa.write_psmcfa(GENOME_FA, seqs)
    return GENOME_FA


def psmc_bin():
    from psmc_b200 import host
    return host.PSMC_BIN


def run_bootstrap(args, fa, total_bins, n_gpus, cb=None):
    """BASELINE metric, second half: wall time of 100 bootstrap replicates (x 25 EM iterations) of the split genome
    (configs[3]; README:49-62 = splitfa + 100 x `psmc -b`), here ONE process: `psmc --split --replicates R --gpus N`
    (host/bootstrap.c: segments resident once per GPU, replicates batched through psmc_b200_set_batch, no collective)."""
    R, iters = args.boot_replicates, args.boot_iters
    base = [psmc_bin(), "-t%g" % MAX_T, "-r%g" % TR_RATIO, "-p", PATTERN, "--split=500000", "--seed", "1", "--gpus", str(n_gpus), "--verbose"]
    # warm-up: a short run of the same binary (device initialisation, page-in of the input file)
    subprocess.run(base + ["-N2", "--replicates", str(2 * n_gpus), "-o", "/tmp/psmc_b200_boot_warm.psmc", fa], capture_output=True, text=True)
    t0 = time.perf_counter()
    r = subprocess.run(base + ["-N%d" % iters, "--replicates", str(R), "-o", "/tmp/psmc_b200_boot.psmc", fa], capture_output=True, text=True)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("psmc --replicates failed: " + r.stderr[-300:])
    tail = [l for l in r.stderr.splitlines() if "bootstrap:" in l]
    em_s = float(tail[-1].split(":")[-1].split()[0]) if tail else None
    done = sum(1 for l in open("/tmp/psmc_b200_boot.psmc") if l.startswith("RD\t%d" % iters))
    out = {"metric": "%d-bootstrap wall-time" % R, "value": wall, "unit": "s", "higher_is_better": False, "n_gpus": n_gpus,
           "replicates": R, "em_iterations": iters, "replicates_completed": done, "em_phase_s": em_s,
           "replicate_iterations_per_s": R * iters / em_s if em_s else None,
           "what": "wall clock of the whole process: reading the .psmcfa text, splitfa rule, upload, %d x %d EM iterations, output; em_phase_s excludes reading" % (R, iters),
           "detail": tail[-1].strip() if tail else None}
    if cb:
        out["cpu_baseline"] = bootstrap_cpu_baseline(cb, total_bins, R, iters)
    return out


def run_decode(args, fa, total_bins, n_gpus):
    """configs[4]: posterior decoding (-d, aux.c:150-182) of the whole genome with fixed parameters (-N0 -i)"""
    par = "/tmp/psmc_b200_bench_params.txt"
    from psmc_b200 import host, synth
    n, nf, _ = host.parse_pattern(PATTERN)
    lam = synth.bottleneck_lambdas(nf)
    with open(par, "w") as fp:      # the layout psmc -i reads (aux.c:84-113): the PA line without its tag
        fp.write("%s %.9f %.9f %.9f %s\n" % (PATTERN, TRUE_THETA, TRUE_RHO, MAX_T, " ".join("%.9f" % x for x in lam)))
    res = {}
    for mode, flag in (("DC", "-d"),):
        cmd = [psmc_bin(), "-N0", "-i", par, flag, "-p", PATTERN, "--gpus", str(n_gpus), "--verbose", "-o", "/tmp/psmc_b200_decode.psmc", fa]
        subprocess.run(cmd, capture_output=True, text=True)        # warm-up
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("psmc -d failed: " + r.stderr[-300:])
        tl = [l for l in r.stderr.splitlines() if "decode:" in l]
        gpu_s = float(tl[-1].split("decode:")[-1].split()[0]) if tl else None
        res[mode] = {"wall_s": wall, "decode_phase_s": gpu_s, "bins_per_s": total_bins / gpu_s if gpu_s else None,
                     "segments": sum(1 for l in open("/tmp/psmc_b200_decode.psmc") if l.startswith("DC")), "detail": tl[-1].strip() if tl else None}
    return {"metric": "decode throughput (-d) on 3Gbp psmcfa, 64 states", "value": res["DC"]["bins_per_s"], "unit": "bins/s", "higher_is_better": True,
            "n_gpus": n_gpus, "bins": total_bins, "modes": res,
            "reference": "oracle/_ref/psmc -N0 -d: 3.65 s per 500 k bins on one core (BASELINE.md section 2) = 1.4e5 bins/s"}


def run_workload_only(args):
    """--workload bootstrap | decode: the secondary metric alone, as its own JSON line (rank 0 of a torchrun launch drives all GPUs)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from psmc_b200 import synth
    seqs = make_genome(args.scale)
    total_bins = sum(len(s) for s in seqs)
    fa = write_genome_file(seqs)
    cores = min(os.cpu_count() or 1, 64)
    if args.workload == "bootstrap":
        cb = cpu_baseline(total_bins, cores, sample_bins=args.cpu_sample_bins, with_c2=False) if not args.no_cpu else None
        o = run_bootstrap(args, fa, total_bins, args.gpus, cb)
        o.update({"steps": 1, "warmup": 1, "ms_per_step": o["value"] * 1e3, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": "configs[3]: %d bootstrap replicates x %d EM iterations of the 22-contig synthetic genome (%d bins) split at 500 000 bins, "
                                         "pattern %s (64 states); one step = the whole job" % (args.boot_replicates, args.boot_iters, total_bins, PATTERN),
                             "parallelism": "replicates dealt to %d GPU(s) in batches, no collective" % args.gpus}})
    else:
        o = run_decode(args, fa, total_bins, args.gpus)
        o.update({"steps": 1, "warmup": 1, "ms_per_step": o["modes"]["DC"]["wall_s"] * 1e3, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                  "config": {"workload": "configs[4]: psmc -N0 -i -d on the 22-contig synthetic genome (%d bins), pattern %s (64 states); one step = the whole job" % (total_bins, PATTERN)}})
    print(json.dumps(o), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="em", choices=["em", "bootstrap", "decode"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale every contig length (1.0 = the 28.8 M-bin workload)")
    ap.add_argument("--chunk", type=int, default=0, help="bins per chunk (0 = auto)")
    ap.add_argument("--cpu-sample-bins", type=int, default=100000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary metrics (bootstrap wall time, decode throughput)")
    ap.add_argument("--boot-replicates", type=int, default=100)
    ap.add_argument("--boot-iters", type=int, default=25)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "em":
        run_workload_only(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
