"""Stage timing of the device pipeline on one simulated sample (device-resident reads).
Usage: python tools/stage_times.py [cfg] [n_pairs]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quasimodo_b200 import Context, _lib, workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
W = {"cfg1": workloads.config1, "cfg2": lambda n: workloads.config2(6, n), "cfg3": workloads.config3,
     "cfg4": workloads.config4, "cfg5": workloads.config5}[cfg](n)
ctx = Context(0)
dev = torch.device("cuda:0")
opt = _lib.default_opt(); opt.w = W.w
idx = ctx.index(W.ref, 31)
L = W.params.read_len
d_genome = torch.from_numpy(W.src_codes).to(dev)
d_codes = torch.empty((2 * n, L), dtype=torch.uint8, device=dev)
d_quals = torch.empty((2 * n, L), dtype=torch.uint8, device=dev)
d_lens = torch.full((2 * n,), L, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ctx.simulate_pairs(W, 0, n, d_genome, d_codes, d_quals, st)
d_regs = torch.zeros(2 * n * _lib.MAX_REGS * 64, dtype=torch.uint8, device=dev)
d_nr = torch.zeros(2 * n, dtype=torch.int32, device=dev)
d_alns = torch.zeros(2 * n * 128, dtype=torch.uint8, device=dev)
d_counts = torch.zeros(_lib.NCH * idx.l_pac, dtype=torch.int32, device=dev)
d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
torch.cuda.synchronize()


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), r


out = {"cfg": W.name, "n_pairs": n}
for rep in range(2):
    d_cells.zero_(); d_counts.zero_()
    t_seed, _ = timed(lambda: ctx.collect_seeds(idx, d_codes, d_lens, st, opt))
    t_se, _ = timed(lambda: ctx.align_se(idx, d_codes, d_lens, d_regs, d_nr, d_cells, st, opt))
    t_pes, pes = timed(lambda: ctx.pestat(idx, d_regs, d_nr, n, st, opt))
    t_pair, _ = timed(lambda: ctx.pair_finish(idx, d_codes, d_lens, d_regs, d_nr, pes, 0, d_alns, st, opt))
    t_pile, _ = timed(lambda: ctx.pileup_accumulate(idx, d_alns, d_codes, d_quals, d_lens, d_counts, st))
    tot = t_se + t_pes + t_pair + t_pile
    out[f"rep{rep}"] = dict(seeds_only_ms=t_seed, align_se_ms=t_se, pestat_ms=t_pes, pair_finish_ms=t_pair, pileup_ms=t_pile,
                            total_ms=tot, pairs_per_s=n / tot * 1e3, cells=int(d_cells.item()),
                            gcups=int(d_cells.item()) / t_se / 1e6)
al = d_alns.cpu().numpy().view(_lib.ALN_DTYPE)
out["mapped_frac"] = float(((al["flag"] & 4) == 0).mean())
out["proper_frac"] = float(((al["flag"] & 2) != 0).mean())
out["depth_mean"] = float(d_counts.view(_lib.NCH, -1)[14].float().mean().item())
print(json.dumps(out, indent=1))
