"""Would two half-batches in flight on one GPU beat one whole batch?  Two contexts (each with its own scratch and streams) on
device 0, two host threads, each aligning + piling up half of a 2 M-pair sample concurrently, against one context doing all of it.
(The insert-size model of the second half is set from the first run, as a multi-GPU rank would get it.)"""
import sys, os, time, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from quasimodo_b200 import Context, _lib, workloads

i = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L, P = 150, 2_000_000
dev = torch.device("cuda:0")
ctxs = [Context(0), Context(0)]
W = workloads.config2(i, P)
opt = _lib.default_opt()
idxs = [c.index(W.ref, 31) for c in ctxs]
smps = [c.sample(ix, opt) for c, ix in zip(ctxs, idxs)]
g = torch.from_numpy(W.src_codes).to(dev)
c = torch.empty((2 * P, L), dtype=torch.uint8, device=dev); q = torch.empty_like(c)
ctxs[0].simulate_pairs(W, 0, P, g, c, q, 0)
lens = torch.full((2 * P,), L, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def whole():
    s = smps[0]
    s.reset(streams[0].cuda_stream)
    s.add_pairs(c, q, lens, pair_id0=0, stream=streams[0].cuda_stream)
    torch.cuda.synchronize()

whole()
pes = smps[0].get_pestat()
ref_counts = smps[0].counts_tensor().clone()

def halves(n_lanes=2):
    def lane(k):
        s = smps[k]
        st = streams[k].cuda_stream
        s.reset(st)
        s.set_pestat(pes)
        lo, hi = k * P // 2, (k + 1) * P // 2
        s.add_pairs(c[2 * lo:2 * hi], q[2 * lo:2 * hi], lens[2 * lo:2 * hi], pair_id0=lo, stream=st)
    th = [threading.Thread(target=lane, args=(k,)) for k in range(2)]
    for t in th: t.start()
    for t in th: t.join()
    torch.cuda.synchronize()

def timed(name, fn, n=5):
    for _ in range(2): fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    print(name, round((time.perf_counter() - t0) / n * 1e3, 2), "ms per 2 M pairs", flush=True)

timed("one context, whole batch", whole)
timed("two contexts, half a batch each, concurrently", halves)
tot = smps[0].counts_tensor() + smps[1].counts_tensor()
print("counts equal:", bool(torch.equal(tot, ref_counts)))
