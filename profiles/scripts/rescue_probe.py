"""Times qm_mate_rescue alone on 1 M pairs of four cfg-2 samples (run from the repo root on a GPU box)."""
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
from quasimodo_b200 import Context, workloads, _lib
ctx = Context(0)
dev = torch.device("cuda:0")
n = 1_000_000
for which in (0, 1, 6, 9):
    W = workloads.config2(which, n)
    idx = ctx.index(W.ref, 31)
    d_codes = None
    if d_codes is None:
        codes, quals, _, _ = W.simulate_host(0, n)
        d_codes = torch.from_numpy(codes).to(dev); d_lens = torch.full((2*n,), W.params.read_len, dtype=torch.int32, device=dev)
    d_regs, d_nr = ctx.align_se(idx, d_codes, d_lens)
    pes = ctx.pestat(idx, d_regs, d_nr, min(n, 65536))
    for rep in range(2):
        r2, n2 = d_regs.clone(), d_nr.clone()
        st = torch.zeros(2, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.mate_rescue(idx, d_codes, d_lens, r2, n2, pes, d_stats=st)
        e1.record(); torch.cuda.synchronize()
    print(which, W.name if hasattr(W, "name") else "", "pes", pes[1], "n_sw", int(st[0]), "cells", int(st[1]), "ms", e0.elapsed_time(e1), "GCUPS", int(st[1]) / e0.elapsed_time(e1) / 1e6, flush=True)
    idx.close()
