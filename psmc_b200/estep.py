"""Python mirror of the C ABI: EStep (context), Model (factored PSMC model)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import CInfo, CModel, CStats, check, load_library

_dp = C.POINTER(C.c_double)


def _d(x):
    return x.ctypes.data_as(_dp)


def factorize(a, tol=1e-9):
    """dense PSMC transition matrix -> dict(U,V,W,Z,D); raises Psmc200Error(ESTRUCT) if it has no such structure."""
    lib = load_library()
    a = np.ascontiguousarray(a, dtype=np.float64)
    N = a.shape[0]
    out = [np.zeros(N) for _ in range(5)]
    check(lib, lib.psmc_b200_factorize(N, _d(a), tol, *[_d(o) for o in out]))
    return dict(zip("UVWZD", out))


class Model:
    """Factored model: a[k][l] = U_k V_l (l<k), W_k Z_l (l>k), D_k; e (2,N); a0 (N,)."""

    def __init__(self, a0, e, U, V, W, Z, D):
        self.a0 = np.ascontiguousarray(a0, dtype=np.float64)
        self.e = np.ascontiguousarray(e, dtype=np.float64).reshape(2, -1)
        self.U, self.V, self.W, self.Z, self.D = [np.ascontiguousarray(x, dtype=np.float64) for x in (U, V, W, Z, D)]
        self.N = len(self.a0)

    @classmethod
    def from_dense(cls, a0, a, e, tol=1e-9):
        f = factorize(a, tol)
        return cls(a0, e, f["U"], f["V"], f["W"], f["Z"], f["D"])

    def dense(self):
        N = self.N
        k, l = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
        return np.where(l < k, np.outer(self.U, self.V), np.where(l > k, np.outer(self.W, self.Z), np.diag(self.D)))

    def c_struct(self):
        m = CModel()
        m.n_states = self.N
        m.a0 = _d(self.a0); m.e = _d(self.e)
        m.U = _d(self.U); m.V = _d(self.V); m.W = _d(self.W); m.Z = _d(self.Z); m.D = _d(self.D)
        return m


class _StatsBuf:
    def __init__(self, N):
        self.N = N
        self.E = np.zeros((2, N))
        self.arr = {k: np.zeros(N) for k in ("RL", "CL", "RU", "CU", "AD")}
        self.c = CStats()
        self.c.E = _d(self.E)
        for k, v in self.arr.items():
            setattr(self.c, k, _d(v))

    def result(self):
        out = dict(LL=self.c.LL, E=self.E.copy())
        out.update({k: v.copy() for k, v in self.arr.items()})
        return out


class EStep:
    """One context per GPU: sequences are uploaded once (2-bit packed) and stay resident.

    Mirrors the reference's em.c:33-55 loop: ``run(model)`` = hmm_pre_backward + for each sequence
    hmm_forward/hmm_backward/hmm_lk/hmm_expect/hmm_add_expect, returning LL and the expected counts.
    """

    def __init__(self, seqs, n_states, device=0, chunk_len=0):
        self.lib = load_library()
        self.N = int(n_states)
        seqs = [np.ascontiguousarray(s, dtype=np.int8) for s in seqs]
        self.L = np.array([len(s) for s in seqs], dtype=np.int32)
        cat = np.ascontiguousarray(np.concatenate(seqs)) if len(seqs) and self.L.sum() else np.zeros(1, dtype=np.int8)
        h = C.c_void_p()
        check(self.lib, self.lib.psmc_b200_create_cat(C.byref(h), len(seqs), self.L.ctypes.data_as(C.POINTER(C.c_int32)),
                                                      cat.ctypes.data_as(C.c_void_p), self.N, device, chunk_len, 0))
        self.h = h
        self.n_seqs = len(seqs)

    def close(self):
        if getattr(self, "h", None):
            self.lib.psmc_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def run(self, model):
        buf = _StatsBuf(self.N)
        m = model.c_struct()
        check(self.lib, self.lib.psmc_b200_estep(self.h, C.byref(m), C.byref(buf.c)))
        return buf.result()

    def run_dense(self, a0, a, e, tol=1e-9):
        buf = _StatsBuf(self.N)
        a0 = np.ascontiguousarray(a0, dtype=np.float64); a = np.ascontiguousarray(a, dtype=np.float64)
        e = np.ascontiguousarray(e, dtype=np.float64)
        check(self.lib, self.lib.psmc_b200_estep_dense(self.h, self.N, _d(a0), _d(a), _d(e), tol, C.byref(buf.c)))
        return buf.result()

    # asynchronous halves (multi-GPU drivers)
    def launch(self, model):
        self._m = model.c_struct()
        self._model_keepalive = model
        check(self.lib, self.lib.psmc_b200_estep_launch(self.h, C.byref(self._m)))

    def device_stats_ptr(self):
        return self.lib.psmc_b200_device_stats(self.h)

    def stats_len(self):
        return self.lib.psmc_b200_stats_len(self.h)

    def stream_ptr(self):
        return self.lib.psmc_b200_stream(self.h)

    def wait(self):
        check(self.lib, self.lib.psmc_b200_wait(self.h))

    def finish(self, n_seqs_total=-1):
        buf = _StatsBuf(self.N)
        check(self.lib, self.lib.psmc_b200_estep_finish(self.h, n_seqs_total, C.byref(buf.c)))
        return buf.result()

    def unpack(self, raw, n_seqs_total):
        buf = _StatsBuf(self.N)
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        check(self.lib, self.lib.psmc_b200_unpack_stats(self.N, _d(raw), n_seqs_total, C.byref(buf.c)))
        return buf.result()

    def decode(self, model, seq_id, full=False, want_s=False):
        """seq_id indexes the records as given at construction (empty records included)"""
        L = int(self.L[seq_id]) if 0 <= seq_id < len(self.L) else 0
        bk = np.zeros(L, dtype=np.int32); bp = np.zeros(L)
        post = np.zeros((L, self.N)) if full else None
        pr = np.zeros(L) if full else None
        s = np.zeros(L) if want_s else None
        m = model.c_struct() if model is not None else None
        check(self.lib, self.lib.psmc_b200_decode(self.h, C.byref(m) if m is not None else None, seq_id,
                                                  bk.ctypes.data_as(C.POINTER(C.c_int32)), _d(bp),
                                                  _d(post) if full else None, _d(pr) if full else None,
                                                  _d(s) if want_s else None))
        return dict(best_k=bk, best_p=bp, post=post, p_recomb=pr, s=s)

    DEC_RUNS, DEC_BINS, DEC_POST = 1, 2, 4

    def decode_all(self, model, runs=True, bins=False, post=False):
        """whole-context decoding on the fast path (psmc_b200_decode_run); returns dict(runs=..., seqs=[per-record dicts])"""
        what = (self.DEC_RUNS if runs else 0) | (self.DEC_BINS if bins else 0) | (self.DEC_POST if post else 0)
        m = model.c_struct()
        check(self.lib, self.lib.psmc_b200_decode_run(self.h, C.byref(m), what))
        out = {}
        if runs:
            n = C.c_int64()
            check(self.lib, self.lib.psmc_b200_decode_get_runs(self.h, 0, None, None, None, None, None, C.byref(n)))
            k = max(int(n.value), 1)
            sq = np.zeros(k, dtype=np.int32); st = np.zeros(k, dtype=np.int32); ln = np.zeros(k, dtype=np.int32)
            ks = np.zeros(k, dtype=np.uint8); mp = np.zeros(k)
            ip = C.POINTER(C.c_int32)
            check(self.lib, self.lib.psmc_b200_decode_get_runs(self.h, k, sq.ctypes.data_as(ip), st.ctypes.data_as(ip), ln.ctypes.data_as(ip),
                                                               ks.ctypes.data_as(C.POINTER(C.c_uint8)), _d(mp), C.byref(n)))
            k = int(n.value)
            out["runs"] = dict(seq=sq[:k], start=st[:k], len=ln[:k], state=ks[:k], max_p=mp[:k])
        if bins or post:
            out["seqs"] = []
            for i, L in enumerate(self.L):
                if L == 0:
                    out["seqs"].append(None)
                    continue
                bk = np.zeros(L, dtype=np.uint8); bp = np.zeros(L, dtype=np.float32)
                po = np.zeros((L, self.N), dtype=np.float32) if post else None
                pr = np.zeros(L) if post else None
                check(self.lib, self.lib.psmc_b200_decode_get_bins(self.h, i, bk.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                                   bp.ctypes.data_as(C.POINTER(C.c_float)),
                                                                   po.ctypes.data_as(C.POINTER(C.c_float)) if post else None,
                                                                   _d(pr) if post else None))
                out["seqs"].append(dict(best_k=bk, best_p=bp, post=po, p_recomb=pr))
        return out

    def set_multiplicity(self, mult=None):
        """bootstrap replicate = multiplicity of every resident record (aux.c:8-47); None restores 1 everywhere"""
        if mult is None:
            check(self.lib, self.lib.psmc_b200_set_multiplicity(self.h, None))
            return
        m = np.ascontiguousarray(mult, dtype=np.int32)
        if m.shape != (self.n_seqs,):
            raise ValueError("mult must have one entry per record given at construction (%d)" % self.n_seqs)
        check(self.lib, self.lib.psmc_b200_set_multiplicity(self.h, m.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_batch(self, mults):
        """batch mode: one model per row of mults (n_rep x n_seqs multiplicities of the resident records)"""
        m = np.ascontiguousarray(mults, dtype=np.int32)
        if m.ndim != 2 or m.shape[1] != self.n_seqs:
            raise ValueError("mults must be (n_rep, %d)" % self.n_seqs)
        check(self.lib, self.lib.psmc_b200_set_batch(self.h, m.shape[0], m.ctypes.data_as(C.POINTER(C.c_int32))))
        self.n_rep = m.shape[0]

    def run_batch(self, models):
        """one E-step of every model of the batch; returns one result dict per model"""
        n = len(models)
        bufs = [_StatsBuf(self.N) for _ in range(n)]
        cm = (CModel * n)(*[m.c_struct() for m in models])
        cs = (CStats * n)(*[b.c for b in bufs])
        check(self.lib, self.lib.psmc_b200_estep_batch(self.h, n, cm, cs))
        out = []
        for i, b in enumerate(bufs):
            b.c = cs[i]
            out.append(b.result())
        return out

    def set_dense(self, on=True):
        """also spill the backward rows in every following E-step (needed by dense_counts)"""
        check(self.lib, self.lib.psmc_b200_set_dense(self.h, 1 if on else 0))

    def dense_counts(self):
        """hmm_expect's dense A[N][N] of the last E-step (khmm.c:305-316), summed over the sequences"""
        A = np.zeros((self.N, self.N))
        check(self.lib, self.lib.psmc_b200_dense_counts(self.h, _d(A)))
        return A

    def set_warm(self, warm_len=-1, eps=0.0):
        """warm-up overlap in bins (0 = exact transfer-matrix path only) and certificate tolerance"""
        check(self.lib, self.lib.psmc_b200_set_warm(self.h, warm_len, eps))

    def info(self):
        inf = CInfo()
        check(self.lib, self.lib.psmc_b200_get_info(self.h, C.byref(inf)))
        d = {f: getattr(inf, f) for f, _ in CInfo._fields_ if f not in ("ms", "decode_ms")}
        d["ms"] = list(inf.ms)
        d["decode_ms"] = list(inf.decode_ms)
        return d
