#!/bin/bash
# round 2, call 3C: the v5 build -- all tests, all five configs, reference arm, launch lists
mkdir -p gpurun_out/r3c
O=gpurun_out/r3c
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
for c in 1 3 4 5; do
  timeout 600 python bench.py --config $c --cpu-seconds 0 > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg5.csv python bench.py --config 5 --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu5.log 2>&1; echo "ncu5 rc=$?"
QM_ROUND_LOG=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 2> $O/roundlog_cfg2.txt > /dev/null
python - <<'PY'
import json
for f in ("default", "cfg1", "cfg3", "cfg4", "cfg5"):
    s = open(f"gpurun_out/r3c/bench_{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
