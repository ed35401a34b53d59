#!/bin/bash
# round 2, call Y: pileup four bases per lane; seed kernel block size
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
for t in 128 64 32; do
  QM_SEED_THREADS=$t timeout 600 python bench.py --cpu-seconds 0 --no-e2e > $O/bench_$t.json 2> $O/bench_$t.err; echo "bench $t rc=$?"
done
python - <<'PY'
import json
for f in ("128", "64", "32"):
    s = open(f"gpurun_out/r2y/bench_{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
