#!/bin/bash
# round 2, call Q: seeding of the host entry's pieces on side streams
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
QM_HOST_TRACE=1 timeout 600 python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > $O/a.json 2> $O/a.err; echo "a rc=$?"
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("a", "bench"):
    s = open(f"gpurun_out/r2q/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["e2e"]["value"])
    print(f, "resident", d["step_ms"], {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
    print(f, "e2e     ", d["e2e"]["step_ms"], {k: round(v, 2) for k, v in d["e2e"]["stages_ms_per_step"].items()})
PY
grep "host trace" $O/a.err | tail -n 2 | cut -c 1-200
timeout 900 python -m pytest tests/test_sample_gpu.py tests/test_driver_gpu.py -m gpu -x -q 2>&1 | tail -n 3
