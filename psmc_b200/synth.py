"""Seeded synthetic .psmcfa-like inputs: sequences drawn from a PSMC HMM (dense a0/a/e
supplied by the caller) with an overlay of missing-data runs.  Pure numpy, deterministic per seed."""
import numpy as np

# human-autosome-like contig lengths in 100-bp bins (chr1..chr22, ~28.8 M bins in total)
HUMAN_AUTOSOME_BINS = [2489564, 2421935, 1982955, 1902145, 1815382, 1708059, 1593459, 1451386, 1383947,
                       1337974, 1350866, 1332753, 1143643, 1070437, 1019911, 903383, 832574, 803732,
                       586176, 644441, 467099, 508184]


def simulate(a0, a, e, L, rng, miss_frac=0.02, miss_mean=50):
    """One sequence of L bins, values 0 (hom) / 1 (het) / 2 (missing), int8."""
    a0 = np.asarray(a0, dtype=np.float64); a = np.asarray(a, dtype=np.float64); e = np.asarray(e, dtype=np.float64)
    N = len(a0)
    stay = np.clip(a.diagonal(), 0.0, 1.0 - 1e-12)
    off = a.copy()
    np.fill_diagonal(off, 0.0)
    off = np.maximum(off, 0.0)
    cum = np.cumsum(off / off.sum(axis=1, keepdims=True), axis=1)
    k = int(np.searchsorted(np.cumsum(a0 / a0.sum()), rng.random()))
    k = min(k, N - 1)
    ks, lens = [], []
    pos = 0
    # batches of random numbers keep the python loop cheap
    while pos < L:
        ub = rng.random(4096); gb = rng.random(4096)
        for u, g in zip(ub, gb):
            d = int(np.log1p(-g) / np.log(stay[k])) + 1 if stay[k] > 0 else 1   # geometric holding time
            ks.append(k); lens.append(d)
            pos += d
            k = min(int(np.searchsorted(cum[k], u)), N - 1)
            if pos >= L:
                break
    states = np.repeat(np.array(ks, dtype=np.int32), np.array(lens, dtype=np.int64))[:L]
    seq = (rng.random(L) < e[1][states]).astype(np.int8)
    if miss_frac > 0:
        n_runs = max(1, int(L * miss_frac / miss_mean))
        starts = rng.integers(0, L, size=n_runs)
        rl = rng.geometric(1.0 / miss_mean, size=n_runs)
        for s_, r_ in zip(starts, rl):
            seq[s_: s_ + r_] = 2
    return seq


def simulate_genome(a0, a, e, lengths, seed, **kw):
    rng = np.random.default_rng(seed)
    return [simulate(a0, a, e, int(L), rng, **kw) for L in lengths]


def iid_track(L, rng, p_het=0.08, p_miss=0.02):
    """cheap non-HMM track for plumbing tests"""
    return rng.choice(np.array([0, 1, 2], dtype=np.int8), size=L, p=[1 - p_het - p_miss, p_het, p_miss])


def bottleneck_lambdas(n_free):
    """a fixed 'true' history with a ~10x bottleneck"""
    x = np.linspace(0, 1, n_free)
    lam = 1.0 + 2.0 * np.exp(-((x - 0.75) / 0.12) ** 2) - 0.9 * np.exp(-((x - 0.35) / 0.08) ** 2)
    return np.maximum(lam, 0.1)
