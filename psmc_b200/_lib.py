"""ctypes loader for libpsmc_b200.so (the C ABI declared in include/psmc_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpsmc_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class LibraryNotBuilt(RuntimeError):
    pass


class Psmc200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("psmc_b200 error %d: %s" % (code, msg))
        self.code = code


class CModel(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("a0", _dp), ("e", _dp),
                ("U", _dp), ("V", _dp), ("W", _dp), ("Z", _dp), ("D", _dp)]


class CStats(C.Structure):
    _fields_ = [("LL", C.c_double), ("E", _dp), ("RL", _dp), ("CL", _dp), ("RU", _dp), ("CU", _dp), ("AD", _dp)]


class CInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_states", C.c_int32), ("n_states_padded", C.c_int32),
                ("n_seqs", C.c_int32), ("n_chunks", C.c_int32), ("chunk_len", C.c_int32),
                ("total_bins", C.c_int64), ("bytes_obs", C.c_int64), ("bytes_forward", C.c_int64),
                ("bytes_transfer", C.c_int64), ("bytes_total", C.c_int64),
                ("ms", C.c_float * 8), ("launches", C.c_int32),
                ("warm_len", C.c_int32), ("fallbacks", C.c_int32), ("fwd_mismatch", C.c_double), ("bwd_mismatch", C.c_double),
                ("failed_fwd", C.c_int32), ("repaired_fwd", C.c_int32), ("failed_bwd", C.c_int32), ("repaired_bwd", C.c_int32),
                ("active_bins", C.c_int64), ("n_seqs_effective", C.c_int64),
                ("n_models", C.c_int32), ("n_chunks_bwd", C.c_int32), ("chunk_len_bwd", C.c_int32),
                ("repair_rounds", C.c_int32), ("warm_redos", C.c_int32), ("decode_ms", C.c_float * 3),
                ("planned", C.c_int32), ("probe_plans", C.c_int32), ("avg_overlap_fwd", C.c_float), ("avg_overlap_bwd", C.c_float),
                ("slow_fwd", C.c_int32), ("slow_bwd", C.c_int32)]


# every symbol include/psmc_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "psmc_b200_version": (C.c_int, []),
    "psmc_b200_last_error": (C.c_char_p, []),
    "psmc_b200_device_count": (C.c_int, []),
    "psmc_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, _ip, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_uint32]),
    "psmc_b200_create_cat": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, _ip, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32]),
    "psmc_b200_destroy": (None, [C.c_void_p]),
    "psmc_b200_upload": (C.c_int, [C.c_void_p, C.c_int32, _ip, C.POINTER(C.c_void_p)]),
    "psmc_b200_upload_cat": (C.c_int, [C.c_void_p, C.c_int32, _ip, C.c_void_p]),
    "psmc_b200_set_multiplicity": (C.c_int, [C.c_void_p, _ip]),
    "psmc_b200_set_batch": (C.c_int, [C.c_void_p, C.c_int32, _ip]),
    "psmc_b200_estep_batch": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(CModel), C.POINTER(CStats)]),
    "psmc_b200_estep_batch_launch": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(CModel)]),
    "psmc_b200_estep_batch_finish": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(CStats)]),
    "psmc_b200_mem_info": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "psmc_b200_estep": (C.c_int, [C.c_void_p, C.POINTER(CModel), C.POINTER(CStats)]),
    "psmc_b200_estep_dense": (C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, C.c_double, C.POINTER(CStats)]),
    "psmc_b200_factorize": (C.c_int, [C.c_int32, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp]),
    "psmc_b200_estep_launch": (C.c_int, [C.c_void_p, C.POINTER(CModel)]),
    "psmc_b200_device_stats": (C.c_void_p, [C.c_void_p]),
    "psmc_b200_stats_len": (C.c_int, [C.c_void_p]),
    "psmc_b200_stream": (C.c_void_p, [C.c_void_p]),
    "psmc_b200_wait": (C.c_int, [C.c_void_p]),
    "psmc_b200_estep_finish": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(CStats)]),
    "psmc_b200_estep_fetch_raw": (C.c_int, [C.c_void_p, _dp]),
    "psmc_b200_unpack_stats": (C.c_int, [C.c_int32, _dp, C.c_int64, C.POINTER(CStats)]),
    "psmc_b200_decode": (C.c_int, [C.c_void_p, C.POINTER(CModel), C.c_int32, _ip, _dp, _dp, _dp, _dp]),
    "psmc_b200_decode_run": (C.c_int, [C.c_void_p, C.POINTER(CModel), C.c_uint32]),
    "psmc_b200_decode_get_runs": (C.c_int, [C.c_void_p, C.c_int64, _ip, _ip, _ip, C.POINTER(C.c_uint8), _dp, C.POINTER(C.c_int64)]),
    "psmc_b200_decode_get_bins": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_float), _dp]),
    "psmc_b200_set_warm": (C.c_int, [C.c_void_p, C.c_int32, C.c_double]),
    "psmc_b200_set_dense": (C.c_int, [C.c_void_p, C.c_int32]),
    "psmc_b200_dense_counts": (C.c_int, [C.c_void_p, _dp]),
    "psmc_b200_get_info": (C.c_int, [C.c_void_p, C.POINTER(CInfo)]),
}

_lib = None


def load_library(path=None):
    """Load libpsmc_b200.so and type every exported symbol.  Fails loudly when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise LibraryNotBuilt("%s not found: build it with `make -C psmc_b200/csrc` (or __graft_entry__.build()); "
                              "there is no CPU fallback" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise Psmc200Error(rc, lib.psmc_b200_last_error().decode(errors="replace"))
