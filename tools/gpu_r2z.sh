#!/bin/bash
# round 2, call Z: pileup with the next pair prefetched into registers
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py tests/test_indels_gpu.py tests/test_depthcap_gpu.py tests/test_dedup_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
QM_SEED_THREADS=64 timeout 600 python bench.py --cpu-seconds 0 --no-e2e > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench",):
    s = open(f"gpurun_out/r2z/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
