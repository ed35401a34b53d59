#!/bin/bash
# round 2, call N: host-entry timeline (QM_HOST_TRACE) and the raw H2D bandwidth of this box
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
python - > $O/h2d.txt 2>&1 <<'PY'
import torch, time
h = torch.empty(600_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device="cuda")
for n in (600_000_000, 57_000_000, 7_000_000):
    for _ in range(2): d[:n].copy_(h[:n], non_blocking=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d[:n].copy_(h[:n], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(n, "B H2D", n * 5 / e0.elapsed_time(e1) / 1e6, "GB/s")
PY
cat $O/h2d.txt
QM_HOST_TRACE=1 timeout 600 python bench.py --cpu-seconds 0 --steps 4 > $O/bench_trace.json 2> $O/bench_trace.err; echo "bench rc=$?"
grep "host trace" $O/bench_trace.err | tail -n 6
