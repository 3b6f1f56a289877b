"""ctypes mirror of libpsmc_host.so (host/): pattern parser, params -> factored model, O(N) objective,
Hooke-Jeeves, the EM session (GPU E-step + host M-step) and the .psmcfa reader."""
import ctypes as C
import os

import numpy as np

from .estep import Model

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(os.path.dirname(_HERE), "host", "libpsmc_host.so")
PSMC_BIN = os.path.join(os.path.dirname(_HERE), "host", "psmc")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lib = None


def _d(x):
    return x.ctypes.data_as(_dp)


def load_host():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RuntimeError("%s not found: build it with `make -C host`" % HOST_LIB_PATH)
        from ._lib import load_library
        load_library()  # libpsmc_b200.so first (the host library links against it)
        L = C.CDLL(HOST_LIB_PATH)
        L.psmch_py_hj.restype = C.c_double
        L.psmch_py_em_create.restype = C.c_void_p
        L.psmch_py_em_ctx.restype = C.c_void_p
        L.psmch_py_read.restype = C.c_void_p
        L.psmch_py_read_name.restype = C.c_char_p
        L.psmch_py_read_seq.restype = C.c_void_p
        L.psmch_py_read_sum.restype = C.c_longlong
        for f in ("psmch_py_em_destroy", "psmch_py_em_iterate", "psmch_py_em_estep", "psmch_py_em_mstep", "psmch_py_em_launch",
                  "psmch_py_read_n", "psmch_py_read_free"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.psmch_py_em_ctx.argtypes = [C.c_void_p, C.c_int]
        L.psmch_py_em_upload.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p]
        L.psmch_py_em_set_raw.argtypes = [C.c_void_p, _dp, C.c_longlong]
        L.psmch_py_em_dims.argtypes = [C.c_void_p, _ip, _ip, _ip]
        L.psmch_py_em_scalars.argtypes = [C.c_void_p, _dp]
        L.psmch_py_em_vectors.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
        L.psmch_py_em_counts.argtypes = [C.c_void_p] + [_dp] * 6
        L.psmch_py_read_sum.argtypes = [C.c_void_p, C.c_int]
        L.psmch_py_read_len.argtypes = [C.c_void_p, C.c_int]
        L.psmch_py_read_name.argtypes = [C.c_void_p, C.c_int]
        L.psmch_py_read_seq.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def parse_pattern(pattern):
    L = load_host()
    nf = C.c_int()
    n = L.psmch_py_pattern(pattern.encode(), C.byref(nf), None)     # first call: sizes only
    if n < 0:
        raise ValueError("bad pattern %r" % pattern)
    pm = np.zeros(n + 1, dtype=np.int32)
    L.psmch_py_pattern(pattern.encode(), C.byref(nf), pm.ctypes.data_as(_ip))
    return n, nf.value, pm


def model_from_params(pattern, params, alpha0=0.1, diverg=False, inp_ti=None):
    """params = [theta, rho, max_t, lambda_free..., (dt)] -> dict with the factored Model and t/sigma/C_pi/C_sigma"""
    L = load_host()
    n, nf, pm = parse_pattern(pattern)
    N = n + 1
    params = np.ascontiguousarray(params, dtype=np.float64)
    t = np.zeros(N + 1); sigma = np.zeros(N); e = np.zeros((2, N)); cc = np.zeros(2)
    F = [np.zeros(N) for _ in range(5)]
    ti = np.ascontiguousarray(inp_ti, dtype=np.float64) if inp_ti is not None else None
    rc = L.psmch_py_model(pattern.encode(), _d(params), C.c_double(alpha0), int(diverg), _d(ti) if ti is not None else None,
                          _d(t), _d(sigma), _d(e), *[_d(x) for x in F], _d(cc))
    if rc < 0:
        raise ValueError("bad pattern %r" % pattern)
    return dict(N=N, n=n, n_free=nf, par_map=pm, t=t, sigma=sigma, a0=sigma, e=e, C_pi=cc[0], C_sigma=cc[1],
                model=Model(sigma, e, *F), params=params.copy(), pattern=pattern)


def avg_t(pattern, params, alpha0=0.1, diverg=False):
    L = load_host()
    n, _, _ = parse_pattern(pattern)
    out = np.zeros(n + 1)
    params = np.ascontiguousarray(params, dtype=np.float64)
    L.psmch_py_avg_t(pattern.encode(), _d(params), C.c_double(alpha0), int(diverg), _d(out))
    return out


def hooke_jeeves(func, x, r=0.5, eps=1e-7, max_calls=50000):
    L = load_host()
    x = np.array(x, dtype=np.float64)
    n = len(x)
    FT = C.CFUNCTYPE(C.c_double, C.c_int, _dp, C.c_void_p)
    calls = [0]

    def cb(n_, xp, _):
        calls[0] += 1
        return float(func(np.ctypeslib.as_array(xp, shape=(n_,))))
    fx = L.psmch_py_hj(FT(cb), n, _d(x), None, C.c_double(r), C.c_double(eps), max_calls)
    return fx, x, calls[0]


def mstep(pattern, params, E, A=None, marg=None, alpha0=0.1):
    """Host M-step on given counts.  Returns dict(params=last evaluated point, Q0, Q1, calls, Q0_offset)."""
    L = load_host()
    params = np.array(params, dtype=np.float64)
    E = np.ascontiguousarray(E, dtype=np.float64)
    res = np.zeros(4)
    if A is not None:
        A = np.ascontiguousarray(A, dtype=np.float64)
        rc = L.psmch_py_mstep(pattern.encode(), C.c_double(alpha0), _d(params), _d(E), _d(A), None, None, None, None, None, _d(res))
    else:
        v = [np.ascontiguousarray(marg[k], dtype=np.float64) for k in ("RL", "CL", "RU", "CU", "AD")]
        rc = L.psmch_py_mstep(pattern.encode(), C.c_double(alpha0), _d(params), _d(E), None, *[_d(x) for x in v], _d(res))
    if rc != 0:
        raise ValueError("mstep failed")
    return dict(params=params, Q0=res[0], Q1=res[1], calls=int(res[2]), Q0_offset=res[3])


class EMSession:
    """The product EM driver (host/em.c): GPU E-step over sharded sequences + host M-step."""

    def __init__(self, pattern, seqs, max_t=15.0, tr_ratio=4.0, alpha0=0.1, init_params=None, devices=(0,), chunk_len=0):
        self.L = load_host()
        seqs = [np.ascontiguousarray(s, dtype=np.int8) for s in seqs]
        lens = np.array([len(s) for s in seqs], dtype=np.int32)
        cat = np.ascontiguousarray(np.concatenate(seqs)) if len(seqs) else np.zeros(1, dtype=np.int8)
        dev = np.array(list(devices), dtype=np.int32)
        ip = np.ascontiguousarray(init_params, dtype=np.float64) if init_params is not None else None
        self.L.psmch_py_em_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.c_void_p, C.c_double, C.c_double,
                                              C.c_double, _dp, C.c_int, C.POINTER(C.c_int32), C.c_int]
        self.h = self.L.psmch_py_em_create(pattern.encode(), len(seqs), lens.ctypes.data_as(C.POINTER(C.c_int32)),
                                           cat.ctypes.data_as(C.c_void_p), max_t, tr_ratio, alpha0,
                                           _d(ip) if ip is not None else None, len(dev), dev.ctypes.data_as(C.POINTER(C.c_int32)), chunk_len)
        if not self.h:
            from ._lib import load_library
            raise RuntimeError("EM session could not be created: %s" % load_library().psmc_b200_last_error().decode())
        n = C.c_int(); nf = C.c_int(); npar = C.c_int()
        self.L.psmch_py_em_dims(self.h, C.byref(n), C.byref(nf), C.byref(npar))
        self.n, self.n_free, self.n_params = n.value, nf.value, npar.value
        self.N = self.n + 1
        self.n_gpus = len(dev)
        self.n_seqs = len(seqs)
        self.pattern = pattern
        self._lens = lens
        self._cat = cat  # host copy of the sequences (for upload())

    def close(self):
        if getattr(self, "h", None):
            self.L.psmch_py_em_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc):
        if rc != 0:
            from ._lib import load_library
            raise RuntimeError("psmc host error %d: %s" % (rc, load_library().psmc_b200_last_error().decode()))

    def upload(self):
        """re-send all sequences host -> device (2-bit pack + H2D into the existing buffers)"""
        self._chk(self.L.psmch_py_em_upload(self.h, self.n_seqs, self._lens.ctypes.data_as(C.POINTER(C.c_int32)),
                                            self._cat.ctypes.data_as(C.c_void_p)))

    def iterate(self):
        self._chk(self.L.psmch_py_em_iterate(self.h))

    def estep(self):
        self._chk(self.L.psmch_py_em_estep(self.h))

    def mstep(self):
        self._chk(self.L.psmch_py_em_mstep(self.h))

    def launch(self):
        self._chk(self.L.psmch_py_em_launch(self.h))

    def ctx(self, g=0):
        return self.L.psmch_py_em_ctx(self.h, g)

    def set_raw(self, raw, n_seqs_total):
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        self._chk(self.L.psmch_py_em_set_raw(self.h, _d(raw), n_seqs_total))

    def set_params(self, params):
        """install parameters (the result of an M-step that ran on another rank); the model is recomputed from them"""
        p = np.ascontiguousarray(params, dtype=np.float64)
        assert p.shape == (self.n_params,)
        self.L.psmch_py_em_set_params.argtypes = [C.c_void_p, _dp]
        self._chk(self.L.psmch_py_em_set_params(self.h, _d(p)))

    def state(self):
        sc = np.zeros(10)
        self.L.psmch_py_em_scalars(self.h, _d(sc))
        params = np.zeros(self.n_params); t = np.zeros(self.N + 1); sigma = np.zeros(self.N); ps = np.zeros(self.N)
        self.L.psmch_py_em_vectors(self.h, _d(params), _d(t), _d(sigma), _d(ps))
        return dict(lk=sc[0], Q0=sc[1], Q1=sc[2], hj_calls=int(sc[3]), t_estep_ms=sc[4], t_mstep_ms=sc[5], C_pi=sc[6],
                    C_sigma=sc[7], sum_L=int(sc[8]), sum_n=int(sc[9]), params=params, t=t, sigma=sigma, post_sigma=ps)

    def counts(self):
        N = self.N
        E = np.zeros((2, N)); v = [np.zeros(N) for _ in range(5)]
        self.L.psmch_py_em_counts(self.h, _d(E), *[_d(x) for x in v])
        return dict(E=E, **dict(zip(("RL", "CL", "RU", "CU", "AD"), v)))


def read_psmcfa(path):
    """(names, seqs) through the product's C parser (host/psmcfa.c)"""
    L = load_host()
    h = L.psmch_py_read(str(path).encode())
    if not h:
        raise IOError("cannot read %s" % path)
    try:
        n = L.psmch_py_read_n(h)
        names, seqs = [], []
        for i in range(n):
            ln = L.psmch_py_read_len(h, i)
            names.append(L.psmch_py_read_name(h, i).decode())
            p = L.psmch_py_read_seq(h, i)
            seqs.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(ln,)).copy() if ln else np.zeros(0, dtype=np.int8))
        return names, seqs, L.psmch_py_read_sum(h, 0), L.psmch_py_read_sum(h, 1)
    finally:
        L.psmch_py_read_free(h)


def split_lengths(lengths, trunk):
    """the splitfa rule (utils/splitfa.c:20-31) as applied by `psmc --split`: [(record, piece number, length)]"""
    L = load_host()
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    cap = int(sum(int(x) // max(trunk, 1) + 2 for x in lengths)) + 4
    oL = np.zeros(cap, dtype=np.int32); rec = np.zeros(cap, dtype=np.int32); idx = np.zeros(cap, dtype=np.int32)
    m = L.psmch_py_split(len(lengths), lengths.ctypes.data_as(_ip), int(trunk), oL.ctypes.data_as(_ip),
                         rec.ctypes.data_as(_ip), idx.ctypes.data_as(_ip), cap)
    if m < 0:
        raise ValueError("bad trunk size")
    return [(int(rec[i]), int(idx[i]), int(oL[i])) for i in range(m)]


def draw_replicate(lengths, seed):
    """one bootstrap replicate as multiplicities (host/bootstrap.c psmch_draw, srand48(seed) stream) and, for
    cross-checking, the multiplicities the copying path of `psmc -b --seed` (host/resamp.c) produces"""
    L = load_host()
    lengths = np.ascontiguousarray(lengths, dtype=np.int32)
    mult = np.zeros(len(lengths), dtype=np.int32); mult2 = np.zeros(len(lengths), dtype=np.int32)
    view = np.zeros(6, dtype=np.int64)
    L.psmch_py_draw.argtypes = [C.c_int, _ip, C.c_long, _ip, C.POINTER(C.c_int64), _ip]
    L.psmch_py_draw(len(lengths), lengths.ctypes.data_as(_ip), int(seed), mult.ctypes.data_as(_ip),
                    view.ctypes.data_as(C.POINTER(C.c_int64)), mult2.ctypes.data_as(_ip))
    return mult, mult2, view
