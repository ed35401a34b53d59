#!/bin/bash
# round 2, call T: pileup with difference arrays
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench",):
    s = open(f"gpurun_out/r2t/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"])
    print(f, {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
    print(d["results"])
PY
