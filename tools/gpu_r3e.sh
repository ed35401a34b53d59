#!/bin/bash
# round 2, call 3E: cfg 4 host entry with one seed launch per landed chunk; cfg 1 with and without the speculative finish
mkdir -p gpurun_out/r3e
O=gpurun_out/r3e
QM_HOST_TRACE=1 timeout 600 python bench.py --config 4 --cpu-seconds 0 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
grep "host trace" $O/bench_cfg4.err | tail -n 2 | cut -c 1-400
QM_ROUND_LOG=1 timeout 600 python bench.py --config 1 --cpu-seconds 0 --no-e2e > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "cfg1 rc=$?"
grep "qm round\|qm spec" $O/bench_cfg1.err | tail -n 4
QM_SPEC=0 timeout 600 python bench.py --config 1 --cpu-seconds 0 --no-e2e > $O/bench_cfg1_nospec.json 2> $O/bench_cfg1_nospec.err; echo "cfg1 rc=$?"
QM_SPEC_MIN=65536 timeout 600 python bench.py --config 1 --cpu-seconds 0 --no-e2e > $O/bench_cfg1_s64k.json 2> $O/bench_cfg1_s64k.err; echo "cfg1 rc=$?"
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench", "cfg1", "cfg1_nospec", "cfg1_s64k", "cfg4"):
    s = open(f"gpurun_out/r3e/bench_{f}.json".replace("bench_bench", "bench")).read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
