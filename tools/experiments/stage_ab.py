"""Stage timers of the resident entry on one 2 M-pair sample of config 2 (TA-1-1 by default): for A/B runs of a kernel
under an environment knob (each knob is read once per process, so one process per setting).
usage: stage_ab.py [sample index 0..9] [label]"""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from quasimodo_b200 import Context, _lib, workloads

i = int(sys.argv[1]) if len(sys.argv) > 1 else 4
label = sys.argv[2] if len(sys.argv) > 2 else ""
L = 150
dev = torch.device("cuda:0")
ctx = Context(0)
P = 2_000_000
W = workloads.config2(i, P)
opt = _lib.default_opt()
idx = ctx.index(W.ref, 31)
s = ctx.sample(idx, opt)
if os.environ.get("QM_AB_BAQ"):
    s.set_baq(int(os.environ["QM_AB_BAQ"]))          # base alignment quality on (3 = as both mpileups run it)
st = torch.cuda.current_stream().cuda_stream
g = torch.from_numpy(W.src_codes).to(dev)
c = torch.empty((2 * P, L), dtype=torch.uint8, device=dev); q = torch.empty_like(c)
ctx.simulate_pairs(W, 0, P, g, c, q, st)
lens = torch.full((2 * P,), L, dtype=torch.int32, device=dev)
torch.cuda.synchronize()

def run():
    s.reset(st)
    s.add_pairs(c, q, lens, pair_id0=0, stream=st)
    torch.cuda.synchronize()

def digest():
    t = s.counts_tensor().to(torch.int64)
    w = (torch.arange(t.shape[1], device=dev) % 1000003 + 1)[None, :] * (torch.arange(t.shape[0], device=dev) * 7919 + 1)[:, None]
    return int((t * w).sum().item())

for _ in range(3): run()
ctx.profile_collect(); ctx.profile_enable(True)
n = 5
t0 = time.perf_counter()
for _ in range(n): run()
dt = (time.perf_counter() - t0) / n * 1e3
ms, _ = ctx.profile_collect(); ctx.profile_enable(False)
print(label, W.name, "ms/step", round(dt, 2), {k: round(v / n, 2) for k, v in ms.items() if v}, "digest", digest(), "stats", s.stats(), flush=True)
