"""Host-side (Python) face of the C-ABI: a thin object wrapper used by tests and bench.py.  Device
memory comes from torch (plumbing only); every computation goes through libquasimodo_b200.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import QmError


def _check(ctx, rc, what):
    if rc != 0:
        msg = _lib.lib().qm_last_error(ctx).decode() if ctx else ""
        raise QmError(f"{what} failed with code {rc}: {msg}")


class Context:
    """One per (process, device).  Raises if no B200-class device is usable (no CPU fallback)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = _lib.lib().qm_ctx_create(int(device), C.byref(self._h))
        if rc != 0:
            self._h = None
            raise QmError(f"qm_ctx_create(device={device}) failed with code {rc}: a CUDA device of compute "
                          "capability 10.x is required; this library has no CPU fallback")
        self.device = device
        self.opt = _lib.default_opt()

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().qm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return _lib.lib().qm_device_sm_count(self._h)

    # ---- extension (ksw_extend2) ----
    def extend_batch_host(self, seq, tasks, opt=None):
        """seq: uint8 arena of base codes; tasks: structured array EXT_TASK_DTYPE -> EXT_RESULT_DTYPE array"""
        opt = opt or self.opt
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        tasks = np.ascontiguousarray(tasks, dtype=_lib.EXT_TASK_DTYPE)
        out = np.zeros(len(tasks), dtype=_lib.EXT_RESULT_DTYPE)
        rc = _lib.lib().qm_extend_batch_host(self._h, C.byref(opt), seq.ctypes.data, seq.nbytes,
                                             tasks.ctypes.data, len(tasks), out.ctypes.data)
        _check(self._h, rc, "qm_extend_batch_host")
        return out

    def extend_batch(self, d_seq, d_tasks, n_tasks, d_out, stream=0, opt=None):
        """device pointers (ints), asynchronous on `stream` (a cudaStream_t handle as int)"""
        opt = opt or self.opt
        rc = _lib.lib().qm_extend_batch(self._h, C.byref(opt), C.c_void_p(d_seq), C.c_void_p(d_tasks),
                                        int(n_tasks), C.c_void_p(d_out), C.c_void_p(stream))
        _check(self._h, rc, "qm_extend_batch")

    def dpx_peak(self, kind=1, iters=4096):
        g, ms = C.c_double(), C.c_double()
        rc = _lib.lib().qm_dpx_peak_sync(self._h, kind, iters, C.byref(g), C.byref(ms))
        _check(self._h, rc, "qm_dpx_peak_sync")
        return g.value, ms.value


def pack_ext_tasks(pairs, h0s, ws, end_bonus, flags=0):
    """Helper: build (seq arena, task array) from a list of (query, target) code arrays."""
    n = len(pairs)
    tasks = np.zeros(n, dtype=_lib.EXT_TASK_DTYPE)
    chunks, off = [], 0
    for i, (q, t) in enumerate(pairs):
        q = np.asarray(q, dtype=np.uint8)
        t = np.asarray(t, dtype=np.uint8)
        tasks[i]["q_off"] = off
        chunks.append(q)
        off += len(q)
        tasks[i]["t_off"] = off
        chunks.append(t)
        off += len(t)
        tasks[i]["qlen"] = len(q)
        tasks[i]["tlen"] = len(t)
    tasks["h0"] = h0s
    tasks["w"] = ws
    tasks["end_bonus"] = end_bonus
    tasks["flags"] = flags
    seq = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint8)
    if len(seq) == 0:
        seq = np.zeros(1, dtype=np.uint8)
    return seq, tasks
