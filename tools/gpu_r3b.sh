#!/bin/bash
# round 2, call 3B: mismatch counting by words in pair_decide; seed blocks of 64
mkdir -p gpurun_out/r3b
O=gpurun_out/r3b
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench",):
    s = open(f"gpurun_out/r3b/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
