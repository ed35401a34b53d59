#!/bin/bash
# round 2, call O: per-sample step times, resident against the host entry
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
QM_HOST_TRACE=1 timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r2o/bench.json").read(); d = json.loads(s[s.index("{"):])
print("resident", d["step_ms"]); print("e2e     ", d["e2e"]["step_ms"])
PY
grep "host trace" $O/bench.err | cut -c 1-200
