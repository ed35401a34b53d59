"""Times qm_mpileup_text (device-resident records -> samtools-mpileup text in device memory) on 2 M pairs of cfg 2 and checks
the text against the count tensor (run from the repo root on a GPU box)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from quasimodo_b200 import Context, workloads, _lib            # noqa: E402
from quasimodo_b200.api import _ptr                             # noqa: E402

ctx = Context(0)
dev = torch.device("cuda:0")
n = 2_000_000
W = workloads.config2(6, n)
idx = ctx.index(W.ref, 31)
codes, quals, _, _ = W.simulate_host(0, n)
d_codes, d_quals = torch.from_numpy(codes).to(dev), torch.from_numpy(quals).to(dev)
d_lens = torch.full((2 * n,), W.params.read_len, dtype=torch.int32, device=dev)
d_regs, d_nr = ctx.align_se(idx, d_codes, d_lens)
pes = ctx.pestat(idx, d_regs, d_nr, 65536)
d_alns = ctx.pair_finish(idx, d_codes, d_lens, d_regs, d_nr, pes)
d_counts = torch.zeros(_lib.NCH * idx.l_pac, dtype=torch.int32, device=dev)
ctx.pileup_accumulate(idx, d_alns, d_codes, d_quals, d_lens, d_counts)
torch.cuda.synchronize()
names = (C.c_char_p * 1)(b"AD169")
for rep in range(3):
    d_text, nbytes = C.c_void_p(), C.c_int64(0)
    t0 = time.perf_counter()
    rc = _lib.lib().qm_mpileup_text(ctx._h, idx._h, C.byref(ctx.pileup_opt), _ptr(d_alns), _ptr(d_codes), _ptr(d_quals), codes.shape[1],
                                    _ptr(d_lens), n, names, C.byref(d_text), C.byref(nbytes), None)
    assert rc == 0
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {nbytes.value / 1e9:.3f} GB of text in {dt * 1e3:.1f} ms = {nbytes.value / dt / 1e9:.1f} GB/s, "
          f"{2 * n / dt / 1e6:.1f} M records/s", flush=True)
# property at full size: the depth column sums to the counted bases plus the deleted positions that pass the quality filter
text = ctx.mpileup_text(idx, d_alns, d_codes, d_quals, d_lens, ["AD169"])
planes = d_counts.cpu().numpy().reshape(_lib.NCH, idx.l_pac)
lines = text.split(b"\n")
depth = sum(int(l.split(b"\t", 4)[3]) for l in lines if l)
counted = int(planes[0:5].sum() + planes[6:11].sum())
dels = int(planes[5].sum() + planes[11].sum())
print("lines", len(lines) - 1, "sum depth", depth, "counted bases", counted, "deleted positions", dels)
assert counted <= depth <= counted + dels
print("ok")
