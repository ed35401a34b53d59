"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/quasimodo_b200.h declares; no compute is called here (there is no GPU in the build box)."""
import ctypes
import os
import re

import pytest

from quasimodo_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so():
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "quasimodo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{]*\{.*?\n\}", "", text, flags=re.S)      # header-only helpers are not exports
    return sorted(set(re.findall(r"\b(qm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(so):
    L = ctypes.CDLL(so)
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/quasimodo_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "quasimodo_b200/_lib.py EXPORTS out of sync with the header"


def test_struct_sizes():
    assert _lib.EXT_TASK_DTYPE.itemsize == 32
    assert _lib.EXT_RESULT_DTYPE.itemsize == 32
    assert ctypes.sizeof(_lib.Opt) == 4 * 17 + 4 * 3 + 4 + 12


def test_defaults_match_bwa_mem_k31(so):
    o = _lib.default_opt()
    assert (o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins) == (1, 4, 6, 1, 6, 1)
    assert (o.w, o.zdrop, o.pen_clip5, o.pen_clip3, o.min_seed_len, o.T) == (100, 100, 5, 5, 31, 30)


def test_no_cpu_fallback(so):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quasimodo_b200 import Context, QmError
    with pytest.raises(QmError):
        Context(0)


def test_missing_nccl_is_an_error_code_not_a_crash(so):
    """NCCL is bound at run time (csrc/comm.cu): on a host without it the library still loads, qm_comm_available() says 0 and the
    communicator entry points answer QM_ENODEV; with the library at hand the same process finds it (fresh processes: the binding
    happens once)"""
    import subprocess
    import sys
    prog = ("import ctypes, sys; L = ctypes.CDLL(sys.argv[1]); buf = (ctypes.c_uint8 * 128)(); "
            "print(L.qm_comm_available(), L.qm_comm_unique_id(buf))")
    env = dict(os.environ, QM_NCCL_LIB="/nonexistent/libnccl.so.2")
    p = subprocess.run([sys.executable, "-c", prog, so], capture_output=True, text=True, env=env)
    assert p.returncode == 0 and p.stdout.split() == ["0", "-2"], (p.stdout, p.stderr)      # QM_ENODEV
    env.pop("QM_NCCL_LIB")
    p = subprocess.run([sys.executable, "-c", "import ctypes, sys; print(ctypes.CDLL(sys.argv[1]).qm_comm_available())", so],
                       capture_output=True, text=True, env=env)
    assert p.returncode == 0 and p.stdout.strip() in ("0", "1")
