#!/bin/bash
# round 2, call M: where the host-entry step spends its time (e2e stage events), chunk-size knob
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 600 python bench.py --cpu-seconds 0 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
QM_HOST_CHUNKS=262144,524288,1048576 timeout 600 python bench.py --cpu-seconds 0 --steps 5 > $O/bench_ramp.json 2> $O/bench_ramp.err; echo "ramp rc=$?"
python - <<'PY'
import json
for f in ("bench_default", "bench_ramp"):
    s = open(f"gpurun_out/r2m/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"])
    print(" resident", {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
    print(" e2e     ", {k: round(v, 2) for k, v in d["e2e"]["stages_ms_per_step"].items()})
PY
