"""Does a host->device copy running beside the kernels slow them down?  Resident step of config 4's sample (2 M pairs) with the
stage timers on: alone, beside a pinned H2D copy loop on another stream, beside a device-to-device copy loop."""
import sys, os, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from quasimodo_b200 import Context, _lib, workloads

cfg = sys.argv[1] if len(sys.argv) > 1 else "4"
P, L = 2_000_000, 150
dev = torch.device("cuda:0")
ctx = Context(0)
W = workloads.config4(50_000_000) if cfg == "4" else workloads.config2(4, P)
opt = _lib.default_opt()
idx = ctx.index(W.ref, 31)
s = ctx.sample(idx, opt)
st = torch.cuda.current_stream().cuda_stream
g = torch.from_numpy(W.src_codes).to(dev)
c = torch.empty((2 * P, L), dtype=torch.uint8, device=dev); q = torch.empty_like(c)
ctx.simulate_pairs(W, 0, P, g, c, q, st)
lens = torch.full((2 * P,), L, dtype=torch.int32, device=dev)
h = torch.empty(600_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty(600_000_000, dtype=torch.uint8, device=dev); d2 = torch.empty_like(d)
side = torch.cuda.Stream()

def run(mode, n=4):
    stop = [False]
    def bg():
        torch.cuda.set_device(0)
        with torch.cuda.stream(side):
            while not stop[0]:
                if mode == "h2d": d.copy_(h, non_blocking=True)
                elif mode == "d2d": d2.copy_(d, non_blocking=True)
                side.synchronize()
    th = None
    if mode != "none":
        th = threading.Thread(target=bg); th.start(); time.sleep(0.05)
    for w in range(2):
        s.reset(st); s.add_pairs(c, q, lens, pair_id0=0, stream=st)
    torch.cuda.current_stream().synchronize()
    ctx.profile_collect(); ctx.profile_enable(True)
    t0 = time.perf_counter()
    for i in range(n):
        s.reset(st); s.add_pairs(c, q, lens, pair_id0=0, stream=st)
    torch.cuda.current_stream().synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    ms, _ = ctx.profile_collect(); ctx.profile_enable(False)
    stop[0] = True
    if th: th.join()
    print(mode, round(dt, 2), {k: round(v / n, 2) for k, v in ms.items() if v})

for m in ("none", "h2d", "d2d", "none"):
    run(m)
