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


def _ptr(t):
    """device pointer of a torch tensor (or None)"""
    return C.c_void_p(0 if t is None else t.data_ptr())


class Index:
    """Device-resident k-mer hash index + reference bases (qm_index)."""

    def __init__(self, ctx, genome, k=31):
        self.ctx, self.genome, self.k = ctx, genome, k
        codes = np.ascontiguousarray(genome.codes, dtype=np.uint8)
        lens = np.ascontiguousarray(genome.lens, dtype=np.int64)
        self._h = C.c_void_p()
        rc = _lib.lib().qm_index_build(ctx._h, codes.ctypes.data, len(lens), lens.ctypes.data, k, C.byref(self._h))
        _check(ctx._h, rc, "qm_index_build")
        self.l_pac = int(_lib.lib().qm_index_lpac(self._h))
        self.offsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            _lib.lib().qm_index_destroy(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
        self.pileup_opt = _lib.default_pileup_opt()

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

    # ---- alignment pipeline on a device-resident read batch (torch uint8 tensors [2*n_pairs, stride]) ----
    def index(self, genome, k=None):
        return Index(self, genome, k or self.opt.min_seed_len)

    def collect_seeds(self, idx, d_codes, d_lens, stream=0, opt=None):
        import torch
        opt = opt or self.opt
        n, stride = d_codes.shape
        seeds = torch.zeros(n * _lib.MAX_SEEDS * 16, dtype=torch.uint8, device=d_codes.device)
        n_seeds = torch.zeros(n, dtype=torch.int32, device=d_codes.device)
        rc = _lib.lib().qm_collect_seeds(self._h, idx._h, C.byref(opt), _ptr(d_codes), stride, _ptr(d_lens), n,
                                         _ptr(seeds), _ptr(n_seeds), C.c_void_p(stream))
        _check(self._h, rc, "qm_collect_seeds")
        return seeds, n_seeds

    def align_se(self, idx, d_codes, d_lens, d_regs=None, d_n_regs=None, d_cells=None, stream=0, opt=None):
        import torch
        opt = opt or self.opt
        n, stride = d_codes.shape
        if d_regs is None:
            d_regs = torch.zeros(n * _lib.MAX_REGS * 64, dtype=torch.uint8, device=d_codes.device)
        if d_n_regs is None:
            d_n_regs = torch.zeros(n, dtype=torch.int32, device=d_codes.device)
        rc = _lib.lib().qm_align_se(self._h, idx._h, C.byref(opt), _ptr(d_codes), stride, _ptr(d_lens), n,
                                    _ptr(d_regs), _ptr(d_n_regs), _ptr(d_cells), C.c_void_p(stream))
        _check(self._h, rc, "qm_align_se")
        return d_regs, d_n_regs

    def pestat(self, idx, d_regs, d_n_regs, n_pairs, stream=0, opt=None):
        opt = opt or self.opt
        pes = np.zeros(4, dtype=_lib.PESTAT_DTYPE)
        rc = _lib.lib().qm_pestat_sync(self._h, idx._h, C.byref(opt), _ptr(d_regs), _ptr(d_n_regs), int(n_pairs),
                                       pes.ctypes.data, C.c_void_p(stream))
        _check(self._h, rc, "qm_pestat_sync")
        return pes

    def pair_finish(self, idx, d_codes, d_lens, d_regs, d_n_regs, pes, pair_id0=0, d_alns=None, stream=0, opt=None):
        import torch
        opt = opt or self.opt
        n, stride = d_codes.shape
        if d_alns is None:
            d_alns = torch.zeros(n * 128, dtype=torch.uint8, device=d_codes.device)
        pes = np.ascontiguousarray(pes, dtype=_lib.PESTAT_DTYPE)
        rc = _lib.lib().qm_pair_finish(self._h, idx._h, C.byref(opt), _ptr(d_codes), stride, _ptr(d_lens), n // 2,
                                       int(pair_id0), _ptr(d_regs), _ptr(d_n_regs), pes.ctypes.data, _ptr(d_alns),
                                       C.c_void_p(stream))
        _check(self._h, rc, "qm_pair_finish")
        return d_alns

    def pileup_accumulate(self, idx, d_alns, d_codes, d_quals, d_lens, d_counts, stream=0, popt=None):
        popt = popt or self.pileup_opt
        n, stride = d_codes.shape
        rc = _lib.lib().qm_pileup_accumulate(self._h, idx._h, C.byref(popt), _ptr(d_alns), _ptr(d_codes), _ptr(d_quals),
                                             stride, _ptr(d_lens), n // 2, _ptr(d_counts), C.c_void_p(stream))
        _check(self._h, rc, "qm_pileup_accumulate")
        return d_counts

    def counts_to_rows(self, idx, d_planes, stream=0):
        import torch
        rows = torch.empty(idx.l_pac * _lib.NCH, dtype=torch.int32, device=d_planes.device)
        rc = _lib.lib().qm_counts_to_rows(self._h, idx._h, _ptr(d_planes), _ptr(rows), C.c_void_p(stream))
        _check(self._h, rc, "qm_counts_to_rows")
        return rows.view(idx.l_pac, _lib.NCH)

    def simulate_pairs(self, workload, pair0, n_pairs, d_genome, d_codes, d_quals, stream=0):
        n, stride = d_codes.shape
        rc = _lib.lib().qm_simulate_pairs(self._h, C.byref(workload.params), _ptr(d_genome), workload.src_off.ctypes.data,
                                          workload.src_len.ctypes.data, workload.src_cum.ctypes.data, int(pair0),
                                          int(n_pairs), int(stride), _ptr(d_codes), _ptr(d_quals), C.c_void_p(stream))
        _check(self._h, rc, "qm_simulate_pairs")


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
