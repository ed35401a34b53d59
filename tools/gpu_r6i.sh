#!/bin/bash
# round 2, call 6I: configs 1, 4 and 5 on the last build
mkdir -p gpurun_out/r6i
O=gpurun_out/r6i
for c in 1 5; do timeout 600 python bench.py --config $c --cpu-seconds 0 > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"; done
timeout 900 python bench.py --config 4 --cpu-seconds 0 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
for f in ("bench_cfg1", "bench_cfg4", "bench_cfg5"):
    s = open(f"gpurun_out/r6i/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
