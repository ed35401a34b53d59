#!/bin/bash
# round 2, call 4A (re-entry): state of HEAD — whole GPU suite, default bench line, launch list
mkdir -p gpurun_out/r4a
O=gpurun_out/r4a
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
QM_ROUND_LOG=1 timeout 300 python bench.py --steps 1 --warmup 1 --cpu-seconds 0 > $O/bench_rl.json 2> $O/roundlog.txt; echo "roundlog rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0 > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("bench",):
    s = open(f"gpurun_out/r4a/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4))
    print(json.dumps(d["stages_ms_per_step"]))
PY
