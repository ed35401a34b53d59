"""Stage timers of the resident entry against the host entry on the same reads (config 4's sample): one chunk, four chunks."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from quasimodo_b200 import Context, _lib, workloads

L = 150
dev = torch.device("cuda:0")
ctx = Context(0)
W = workloads.config4(50_000_000)
opt = _lib.default_opt()
idx = ctx.index(W.ref, 31)
s = ctx.sample(idx, opt)
st = torch.cuda.current_stream().cuda_stream
g = torch.from_numpy(W.src_codes).to(dev)
CH = 1 << 21
P = 4 * CH
c = torch.empty((2 * P, L), dtype=torch.uint8, device=dev); q = torch.empty_like(c)
for o in range(0, P, CH):
    ctx.simulate_pairs(W, o, CH, g, c[2 * o:2 * (o + CH)], q[2 * o:2 * (o + CH)], st)
lens = torch.full((2 * P,), L, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
h_c = c.cpu().pin_memory(); h_q = q.cpu().pin_memory(); h_l = lens.cpu().pin_memory()
from quasimodo_b200.api import pack_reads
h_b2 = torch.empty((2 * P, (L + 3) // 4), dtype=torch.uint8).pin_memory(); h_nm = torch.empty((2 * P, (L + 7) // 8), dtype=torch.uint8).pin_memory()
b2, nm = pack_reads(h_c.numpy())
h_b2.copy_(torch.from_numpy(b2)); h_nm.copy_(torch.from_numpy(nm))

def timed(name, fn, n=3):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ctx.profile_collect(); ctx.profile_enable(True)
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    ms, _ = ctx.profile_collect(); ctx.profile_enable(False)
    print(name, round(dt, 2), {k: round(v / n, 2) for k, v in ms.items() if v}, flush=True)

def resident(n_chunks):
    def f():
        s.reset(st)
        for k in range(n_chunks):
            s.add_pairs(c[2 * k * CH:2 * (k + 1) * CH], q[2 * k * CH:2 * (k + 1) * CH], lens[:2 * CH], pair_id0=k * CH, stream=st)
        torch.cuda.synchronize()
    return f
def host(n_chunks, packed):
    def f():
        s.reset(st); torch.cuda.synchronize()
        n = n_chunks * CH
        if packed: s.add_pairs_host_packed(h_b2[:2 * n], h_nm[:2 * n], h_q[:2 * n], h_l[:2 * n], pair_id0=0)
        else: s.add_pairs_host(h_c[:2 * n], h_q[:2 * n], h_l[:2 * n], pair_id0=0)
    return f
timed("resident x1", resident(1)); timed("host packed x1", host(1, True)); timed("host bytes x1", host(1, False))
timed("resident x4", resident(4)); timed("host packed x4", host(4, True)); timed("host bytes x4", host(4, False))
