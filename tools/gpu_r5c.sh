#!/bin/bash
# round 2, call 5C: eight GPUs, the default bench line (config 2, weak scaling) and config 4 (strong scaling), end-of-round build
mkdir -p gpurun_out/r5c
O=gpurun_out/r5c
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --cpu-seconds 0 > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "8gpu rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 4 --cpu-seconds 0 --steps 3 --warmup 1 > $O/bench_8gpu_cfg4.json 2> $O/bench_8gpu_cfg4.err; echo "8gpu cfg4 rc=$?"
python - <<'PY'
import json
for f in ("bench_8gpu", "bench_8gpu_cfg4"):
    try:
        s = open(f"gpurun_out/r5c/{f}.json").read(); d = json.loads(s[s.index("{"):])
        print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), d["n_gpus"], d["scaling"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -n 3 $O/bench_8gpu.err | cut -c1-300
