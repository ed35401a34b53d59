#!/bin/bash
# round 2, call R: tail kernel with column slots by query length; CIGAR score pass with the band narrowed by cig_gain_cap
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --cpu-seconds 0 --config 5 --no-e2e > $O/bench5.json 2> $O/bench5.err; echo "bench5 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench5"):
    s = open(f"gpurun_out/r2r/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"])
    print(f, {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
