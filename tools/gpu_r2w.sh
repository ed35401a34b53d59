#!/bin/bash
# round 2, call W: threshold of the speculative finish
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
for m in 131072 65536 32768 16384; do
  QM_SPEC_MIN=$m timeout 600 python bench.py --cpu-seconds 0 --no-e2e > $O/bench_$m.json 2> $O/bench_$m.err; echo "bench $m rc=$?"
done
QM_SPEC_MIN=131072 QM_ROUND_LOG=1 timeout 600 python bench.py --cpu-seconds 0 --steps 2 --warmup 3 --no-e2e > $O/bench_log.json 2> $O/bench_log.err
grep "qm spec" $O/bench_log.err | tail -n 5
python - <<'PY'
import json
for f in ("131072", "65536", "32768", "16384"):
    s = open(f"gpurun_out/r2w/bench_{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items() if k in ("advance", "extend", "other")})
PY
