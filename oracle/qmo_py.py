"""ctypes binding of the CPU oracle (oracle/libqmo.so).  TEST INFRASTRUCTURE ONLY: importable from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from the
product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Opt(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "zdrop", "pen_clip5", "pen_clip3",
                 "min_seed_len", "max_occ", "T", "pen_unpaired", "max_ins", "max_chain_gap", "mapq_coef_len")] + \
               [("mask_level", C.c_float), ("drop_ratio", C.c_float), ("mask_level_redun", C.c_float),
                ("min_chain_weight", C.c_int32), ("reserved", C.c_int32 * 3)]


EXT_DTYPE = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"),
                      ("gscore", "<i4"), ("max_off", "<i4")])


def build(force=False):
    so = os.path.join(_HERE, "libqmo.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.qmo_ksw_extend2.restype = C.c_int64
        _LIB.qmo_ksw_global2.restype = C.c_int
    return _LIB


def default_opt():
    o = Opt()
    lib().qmo_opt_default(C.byref(o))
    return o


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


def ksw_extend2(query, target, h0, w, end_bonus, opt=None):
    """-> ((score,qle,tle,gtle,gscore,max_off), executed_cells)"""
    opt = opt or default_opt()
    q, qp = _u8(query)
    t, tp = _u8(target)
    out = np.zeros(1, dtype=EXT_DTYPE)
    cells = lib().qmo_ksw_extend2(len(q), qp, len(t), tp, C.byref(opt), int(w), int(end_bonus), int(h0),
                                  out.ctypes.data_as(C.c_void_p))
    return tuple(int(x) for x in out[0]), int(cells)


def ksw_global2(query, target, w, opt=None, max_cigar=64):
    """-> (score, [(op,len)...]) with op 0=M 1=I 2=D"""
    opt = opt or default_opt()
    q, qp = _u8(query)
    t, tp = _u8(target)
    cig = np.zeros(max_cigar, dtype=np.uint32)
    n = C.c_int(0)
    s = lib().qmo_ksw_global2(len(q), qp, len(t), tp, C.byref(opt), int(w), C.byref(n),
                              cig.ctypes.data_as(C.c_void_p), max_cigar)
    if n.value < 0:
        raise OverflowError("cigar overflow")
    return int(s), [(int(c & 0xf), int(c >> 4)) for c in cig[:n.value]]
