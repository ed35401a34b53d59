"""CPU parity of the base-alignment-quality kernels' logic (quasimodo_b200/csrc/baq_core.cuh, the statements baq_fast_kernel and
baq_slow_kernel run) compiled for the host over a bounds-checked, poisoned scratch (tests/baq_host.cpp) against the oracle's
restatement of htslib's sam_prob_realn / kpa_glocal (oracle/qmo_baq.c): the capped qualities byte for byte, plain and extended BAQ,
wide bands included.  The GPU test (tests/test_baq_gpu.py) then shows the kernels agree with the same oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import qmo_py

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host():
    bdir = os.path.join(HERE, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libbaqhost.so")
    srcs = [os.path.join(HERE, "baq_host.cpp"), os.path.join(ROOT, "quasimodo_b200", "csrc", "baq_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.baq_host.restype = C.c_longlong
    return lib


@pytest.mark.parametrize("flag", [3, 1])
def test_kernel_logic_on_the_host_equals_the_oracle(host, flag):
    from quasimodo_b200 import workloads
    n, L = 1500, 250
    W = workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    codes = codes.copy()
    ref = W.ref.codes
    rng = np.random.default_rng(7)
    for i in range(30):                                    # reads with a 15-base deletion: the wide-band class
        p = int(rng.integers(1000, len(ref) // 2))
        codes[2 * i] = np.concatenate([ref[p:p + 120], ref[p + 135:p + 135 + L - 120]])
        codes[2 * i + 1] = (3 - ref[p + 300:p + 300 + L])[::-1]
    lens = np.full(2 * n, L, np.int32)
    opt = qmo_py.default_opt()
    opt.w = 200
    oref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, _, _, _ = qmo_py.run_sample(oref, codes, quals, lens, opt=opt)
    want = qmo_py.baq(oref, alns, codes, quals, lens, flag=flag)
    popt = qmo_py.PileupOpt()
    qmo_py.lib().qmo_pileup_opt_default(C.byref(popt))
    got = np.empty_like(quals)
    off = np.concatenate([[0], np.cumsum(W.ref.lens)[:-1]]).astype(np.int64)
    ln = np.asarray(W.ref.lens, np.int64)
    rc = np.ascontiguousarray(W.ref.codes, np.uint8)
    a = np.ascontiguousarray(alns)
    wide = host.baq_host(rc.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), ln.ctypes.data_as(C.c_void_p), C.byref(popt),
                         a.ctypes.data_as(C.c_void_p), codes.ctypes.data_as(C.c_void_p), quals.ctypes.data_as(C.c_void_p), C.c_int(L),
                         lens.ctypes.data_as(C.c_void_p), C.c_longlong(2 * n), C.c_int(flag), got.ctypes.data_as(C.c_void_p))
    assert wide >= 15
    assert (want < quals).any()
    bad = np.argwhere(got != want)
    assert len(bad) == 0, (len(bad), bad[:5])
