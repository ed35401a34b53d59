"""Micro-benchmark of the extension kernel alone (device-resident tasks) + DPX issue-rate peak.
Usage: python tools/bench_extend.py [n_tasks]"""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from quasimodo_b200 import Context, _lib
from quasimodo_b200.api import pack_ext_tasks
from tests import extgen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
ctx = Context(0)
out = {}
for kind, name in [(0, "viaddmax_s32_relu"), (1, "viaddmax_s16x2_relu"), (2, "vimax3_s32(+add)")]:
    g, ms = ctx.dpx_peak(kind, 8192)
    out["dpx_" + name] = {"gops_lane": g, "ms": ms}
rng = np.random.default_rng(1)
# read-like tasks: query = read tail (1..119), target = mutated copy + flank, h0 = seed score
base_pairs, h0s, ws = [], [], []
m = 4096
for _ in range(m):
    ql = int(rng.integers(20, 120))
    q = rng.integers(0, 4, ql).astype(np.uint8)
    t = extgen.mutate(rng, q, 0.04, 0.004)
    t = np.concatenate([t, rng.integers(0, 4, ql - 5 if ql > 6 else 1).astype(np.uint8)])
    base_pairs.append((q, t)); h0s.append(int(rng.integers(31, 120))); ws.append(100)
seq, tasks = pack_ext_tasks(base_pairs, h0s, ws, 5, 1)
reps = (n + m - 1) // m
tasks_big = np.tile(tasks, reps)[:n]
if len(sys.argv) > 2 and sys.argv[2] == "sorted":      # warp-mates of similar shape: sensitivity of the thread-per-task kernel
    tasks_big = tasks_big[np.lexsort((tasks_big["h0"], tasks_big["qlen"]))]
d_seq = torch.from_numpy(seq).cuda()
d_tasks = torch.from_numpy(tasks_big.view(np.uint8)).cuda()
d_out = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.extend_batch(d_seq.data_ptr(), d_tasks.data_ptr(), n, d_out.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for _ in range(K):
    ctx.extend_batch(d_seq.data_ptr(), d_tasks.data_ptr(), n, d_out.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
res = d_out.cpu().numpy().view(_lib.EXT_RESULT_DTYPE)
cells = int(res["cells"].astype(np.int64).sum())
nominal = int((tasks_big["qlen"].astype(np.int64) * tasks_big["tlen"]).sum())
out["extend"] = {"n_tasks": n, "ms": ms, "cells": cells, "gcups_executed": cells / ms / 1e6,
                 "gcups_nominal": nominal / ms / 1e6, "tasks_per_s": n / ms * 1e3}
print(json.dumps(out, indent=1))
