#!/bin/bash
# round 2, call 5H: where the file-level time goes (the driver's phase clock), 4 M pairs
mkdir -p gpurun_out/r5h
O=gpurun_out/r5h
timeout 900 python tools/file_level_bench.py 4000000 4 16 > $O/file_level_4m.jsonl 2> $O/file_level_4m.err; echo "file-level rc=$?"
python - <<'PY'
import json
for ln in open("gpurun_out/r5h/file_level_4m.jsonl"):
    d = json.loads(ln); print(d["outputs"], d["wall_s"], d["value"]); print("   ", d["phases"])
PY
