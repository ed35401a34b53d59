#!/bin/bash
# round 2, call V: speculative finish of a batch's last reads
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_extend_gpu.py -m gpu -x -q > $O/pytest1.log 2>&1; echo "pytest1 rc=$?"; tail -n 15 $O/pytest1.log
QM_ROUND_LOG=1 timeout 600 python bench.py --cpu-seconds 0 --steps 2 --warmup 3 --no-e2e > $O/bench_log.json 2> $O/bench_log.err; echo "bench rc=$?"
grep "qm spec\|qm round" $O/bench_log.err | tail -n 8
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
QM_SPEC=0 timeout 600 python bench.py --cpu-seconds 0 --no-e2e > $O/bench_nospec.json 2> $O/bench_nospec.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_nospec"):
    s = open(f"gpurun_out/r2v/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"])
    print(f, {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
