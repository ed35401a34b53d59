#!/bin/bash
# round 2, call 6A: the last build of the round -- whole GPU suite, smoke, the default bench line and its launch list
mkdir -p gpurun_out/r6a
O=gpurun_out/r6a
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r6a/bench_default.json").read(); d = json.loads(s[s.index("{"):])
print(round(d["value"] / 1e6, 3), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), d["roofline"]["traffic"], d["cpu_baseline"]["value"], d["gpu_launches"])
print({k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
