"""ctypes bindings for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Two checkers live here:
  * ``Oracle``  : oracle/liboracle.so, our plain-C restatement of the reference (psmc_oracle.c).
  * ``Ref``     : oracle/_ref/libpsmcref.so, the UNMODIFIED reference objects behind flat-array
                  entry points (ref_harness.c).  Present whenever oracle/_ref was built (it is built
                  in the dev container from /root/reference and shipped prebuilt to the GPU box).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product (psmc_b200/, host/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference exists, oracle/_ref/*."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _d(x):
    return x.ctypes.data_as(_dp)


def _i(x):
    return x.ctypes.data_as(_ip)


def _seq_bytes(seqs):
    """list of int8 arrays (values 0/1/2) -> (L int32 array, concatenated int8 array)"""
    L = np.array([len(s) for s in seqs], dtype=np.int32)
    cat = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int8) for s in seqs]))
    return L, cat


class _Base:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)


class Oracle(_Base):
    """Our CPU restatement (kind == "port")."""
    prefix = "orc_"
    kind = "port"

    def __init__(self):
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        super().__init__(path)
        L = self.lib
        L.orc_lk.restype = C.c_double
        L.orc_Q0.restype = C.c_double
        L.orc_Q.restype = C.c_double
        L.orc_hj.restype = C.c_double

    def pattern(self, pattern):
        nf = C.c_int()
        pm = np.zeros(1024, dtype=np.int32)
        n = self.lib.orc_pattern(pattern.encode(), C.byref(nf), _i(pm))
        if n < 0:
            raise ValueError("bad pattern %r" % pattern)
        return n, nf.value, pm[: n + 1].copy()

    def update_hmm(self, pattern, params, alpha=0.1, diverg=False, inp_ti=None):
        n, nf, pm = self.pattern(pattern)
        N = n + 1
        params = np.ascontiguousarray(params, dtype=np.float64)
        a = np.zeros((N, N)); e = np.zeros((2, N)); a0 = np.zeros(N); sigma = np.zeros(N); t = np.zeros(n + 2)
        cpi = C.c_double(); csig = C.c_double()
        ti = None
        if inp_ti is not None:
            ti = np.ascontiguousarray(inp_ti, dtype=np.float64)
        self.lib.orc_update_hmm(n, _i(pm), _d(params), len(params), C.c_double(alpha), int(diverg),
                                _d(ti) if ti is not None else None,
                                _d(a), _d(e), _d(a0), _d(sigma), _d(t), C.byref(cpi), C.byref(csig))
        return dict(N=N, n=n, n_free=nf, par_map=pm, a=a, e=e, a0=a0, sigma=sigma, t=t,
                    C_pi=cpi.value, C_sigma=csig.value)

    def avg_t(self, pattern, params, model, diverg=False):
        n, nf, pm = self.pattern(pattern)
        params = np.ascontiguousarray(params, dtype=np.float64)
        out = np.zeros(n + 1)
        self.lib.orc_avg_t(n, _i(pm), _d(params), len(params), int(diverg), _d(model["t"]), _d(model["sigma"]),
                           C.c_double(model["C_pi"]), C.c_double(model["C_sigma"]), _d(out))
        return out

    def fwdbwd(self, a, e, a0, seq):
        N = len(a0); L = len(seq)
        a = np.ascontiguousarray(a); e = np.ascontiguousarray(e); a0 = np.ascontiguousarray(a0)
        seq = np.ascontiguousarray(seq, dtype=np.int8)
        f = np.zeros((L, N)); b = np.zeros((L, N)); s = np.zeros(L)
        self._fn("fwdbwd")(N, _d(a), _d(e), _d(a0), L, seq.ctypes.data_as(C.c_char_p), _d(f), _d(b), _d(s))
        return f, b, s

    def estep(self, a, e, a0, seqs):
        N = len(a0)
        a = np.ascontiguousarray(a); e = np.ascontiguousarray(e); a0 = np.ascontiguousarray(a0)
        L, cat = _seq_bytes(seqs)
        LL = C.c_double()
        A = np.zeros((N, N)); E = np.zeros((2, N)); A0 = np.zeros(N)
        self.lib.orc_estep(N, _d(a), _d(e), _d(a0), len(L), _i(L), cat.ctypes.data_as(C.c_char_p),
                           C.byref(LL), _d(A), _d(E), _d(A0))
        return dict(LL=LL.value, A=A, E=E, A0=A0)

    def decode(self, a, e, a0, seq, full=True):
        N = len(a0); L = len(seq)
        a = np.ascontiguousarray(a); e = np.ascontiguousarray(e); a0 = np.ascontiguousarray(a0)
        seq = np.ascontiguousarray(seq, dtype=np.int8)
        bk = np.zeros(L, dtype=np.int32); bp = np.zeros(L)
        post = np.zeros((L, N)) if full else None
        pr = np.zeros(L) if full else None
        self._fn("decode")(N, _d(a), _d(e), _d(a0), L, seq.ctypes.data_as(C.c_char_p), _i(bk), _d(bp),
                           _d(post) if full else None, _d(pr) if full else None)
        return dict(best_k=bk, best_p=bp, post=post, p_recomb=pr)

    def Q0(self, A, E):
        N = A.shape[0]
        return self.lib.orc_Q0(N, _d(np.ascontiguousarray(A)), _d(np.ascontiguousarray(E)))

    def Q(self, a, e, A, E, Q0):
        N = A.shape[0]
        return self.lib.orc_Q(N, _d(np.ascontiguousarray(a)), _d(np.ascontiguousarray(e)),
                              _d(np.ascontiguousarray(A)), _d(np.ascontiguousarray(E)), C.c_double(Q0))

    def hj(self, func, x, r=0.5, eps=1e-7, max_calls=50000):
        x = np.array(x, dtype=np.float64)
        n = len(x)
        calls = [0]
        FT = C.CFUNCTYPE(C.c_double, C.c_int, _dp, C.c_void_p)

        def cb(n_, xp, _):
            calls[0] += 1
            return float(func(np.ctypeslib.as_array(xp, shape=(n_,))))
        fx = self._fn("hj" if self.prefix == "orc_" else "kmin_hj")(FT(cb), n, _d(x), None, C.c_double(r),
                                                                      C.c_double(eps), max_calls)
        return fx, x, calls[0]

    def struct_stats(self, A):
        N = A.shape[0]
        out = [np.zeros(N) for _ in range(5)]
        self.lib.orc_struct_stats(N, _d(np.ascontiguousarray(A)), *[_d(o) for o in out])
        return dict(zip(["RL", "CL", "RU", "CU", "AD"], out))

    def factors(self, a):
        N = a.shape[0]
        out = [np.zeros(N) for _ in range(5)]
        self.lib.orc_factors(N, _d(np.ascontiguousarray(a)), *[_d(o) for o in out])
        return dict(zip(["U", "V", "W", "Z", "D"], out))


class Ref(Oracle):
    """The unmodified reference behind ref_harness.c (kind == "reference")."""
    prefix = "ref_"
    kind = "reference"

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libpsmcref.so")
        if not os.path.exists(path) and os.path.exists("/root/reference/khmm.c"):
            build()
        _Base.__init__(self, path)
        self.lib.ref_Q.restype = C.c_double
        self.lib.ref_kmin_hj.restype = C.c_double
        self.psmc_bin = os.path.join(_HERE, "_ref", "psmc")
        self.splitfa_bin = os.path.join(_HERE, "_ref", "splitfa")

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libpsmcref.so"))

    def pattern(self, pattern):
        nf = C.c_int()
        pm = np.zeros(1024, dtype=np.int32)
        n = self.lib.ref_pattern(pattern.encode(), C.byref(nf), _i(pm))
        return n, nf.value, pm[: n + 1].copy()

    def update_hmm(self, pattern, params, alpha=0.1, diverg=False, inp_ti=None):
        assert inp_ti is None
        n, nf, pm = self.pattern(pattern)
        N = n + 1
        params = np.ascontiguousarray(params, dtype=np.float64)
        a = np.zeros((N, N)); e = np.zeros((2, N)); a0 = np.zeros(N); sigma = np.zeros(N); t = np.zeros(n + 2)
        cpi = C.c_double(); csig = C.c_double()
        self.lib.ref_update_hmm(pattern.encode(), _d(params), C.c_double(alpha), int(diverg),
                                _d(a), _d(e), _d(a0), _d(sigma), _d(t), C.byref(cpi), C.byref(csig))
        t[n + 1] = 1000.0
        return dict(N=N, n=n, n_free=nf, par_map=pm, a=a, e=e, a0=a0, sigma=sigma, t=t,
                    C_pi=cpi.value, C_sigma=csig.value)

    def estep(self, a, e, a0, seqs):
        N = len(a0)
        a = np.ascontiguousarray(a); e = np.ascontiguousarray(e); a0 = np.ascontiguousarray(a0)
        L, cat = _seq_bytes(seqs)
        LL = C.c_double(); Q0 = C.c_double()
        A = np.zeros((N, N)); E = np.zeros((2, N)); A0 = np.zeros(N)
        self.lib.ref_estep(N, _d(a), _d(e), _d(a0), len(L), _i(L), cat.ctypes.data_as(C.c_char_p),
                           C.byref(LL), _d(A), _d(E), _d(A0), C.byref(Q0))
        return dict(LL=LL.value, A=A, E=E, A0=A0, Q0=Q0.value)

    def Q0(self, A, E):
        raise NotImplementedError("use Q_ref")

    def Q_ref(self, a, e, a0, A, E):
        N = A.shape[0]
        q0 = C.c_double()
        q = self.lib.ref_Q(N, _d(np.ascontiguousarray(a)), _d(np.ascontiguousarray(e)), _d(np.ascontiguousarray(a0)),
                           _d(np.ascontiguousarray(A)), _d(np.ascontiguousarray(E)), C.byref(q0))
        return q, q0.value
