#!/bin/bash
# round 2, call 5B: extension refill fetches the query eight bases per trip; bench e2e with the records returned; truth sets filtered
mkdir -p gpurun_out/r5b
O=gpurun_out/r5b
timeout 900 python -m pytest tests/test_extend_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 300 python tools/experiments/stage_ab.py 4 "TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 9 "TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r5b/bench.json").read(); d = json.loads(s[s.index("{"):])
print(round(d["value"] / 1e6, 3), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), d["e2e"].get("with_records"))
print({k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
print(d["results"])
PY
