"""ctypes loader of the C-ABI library (quasimodo_b200/libquasimodo_b200.so).

There is no CPU fallback: if the library is missing this raises, and creating a context without a
B200 raises.  The struct layouts mirror include/quasimodo_b200.h."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libquasimodo_b200.so")

_LIB = None


class QmError(RuntimeError):
    pass


class Opt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "zdrop", "pen_clip5", "pen_clip3",
                 "min_seed_len", "max_occ", "T", "pen_unpaired", "max_ins", "max_chain_gap", "mapq_coef_len")] + \
               [("mask_level", C.c_float), ("drop_ratio", C.c_float), ("mask_level_redun", C.c_float),
                ("min_chain_weight", C.c_int32), ("reserved", C.c_int32 * 3)]


class SimParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("read_len", C.c_int32), ("ins_mean", C.c_int32), ("ins_sd", C.c_int32),
                ("ins_max", C.c_int32), ("n_sources", C.c_int32), ("indel_ppm", C.c_int32), ("n_ppm", C.c_int32),
                ("lowq_ppm", C.c_int32), ("reserved", C.c_int32 * 3)]


EXT_TASK_DTYPE = np.dtype([("q_off", "<u4"), ("t_off", "<u4"), ("qlen", "<i4"), ("tlen", "<i4"),
                           ("h0", "<i4"), ("w", "<i4"), ("end_bonus", "<i4"), ("flags", "<u4")])
EXT_RESULT_DTYPE = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                             ("gscore", "<i4"), ("max_off", "<i4"), ("w_used", "<i4"), ("cells", "<i4")])
QM_EXT_BAND_RETRY = 1
QM_EXT_PREV_H0 = 2

# every symbol include/quasimodo_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = [
    "qm_opt_default", "qm_ctx_create", "qm_ctx_destroy", "qm_last_error", "qm_version",
    "qm_device_sm_count", "qm_extend_batch", "qm_extend_batch_host", "qm_dpx_peak_sync",
    "qm_simulate_pairs_host", "qm_simulate_pairs",
]


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise QmError(f"{SO_PATH} is missing: build it with `python -m quasimodo_b200.build` "
                          "(there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        L.qm_version.restype = C.c_char_p
        L.qm_last_error.restype = C.c_char_p
        L.qm_last_error.argtypes = [C.c_void_p]
        L.qm_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.qm_ctx_destroy.argtypes = [C.c_void_p]
        L.qm_ctx_destroy.restype = None
        L.qm_device_sm_count.argtypes = [C.c_void_p]
        L.qm_extend_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.qm_extend_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64, C.c_void_p]
        L.qm_dpx_peak_sync.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.qm_simulate_pairs_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qm_simulate_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def default_opt():
    o = Opt()
    lib().qm_opt_default(C.byref(o))
    return o
