#!/bin/bash
# round 2, call P: one sample, stage by stage: resident, host entry, host entry without copy/compute overlap
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
QM_HOST_TRACE=1 timeout 600 python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > $O/a.json 2> $O/a.err; echo "a rc=$?"
QM_HOST_SEQ=1 QM_HOST_TRACE=1 timeout 600 python bench.py --cpu-seconds 0 --steps 1 --warmup 3 > $O/b.json 2> $O/b.err; echo "b rc=$?"
python - <<'PY'
import json
for f in "ab":
    s = open(f"gpurun_out/r2p/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, "resident", d["step_ms"], {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
    print(f, "e2e     ", d["e2e"]["step_ms"], {k: round(v, 2) for k, v in d["e2e"]["stages_ms_per_step"].items()})
PY
grep "host trace" $O/a.err | tail -n 2 | cut -c 1-200; grep "host trace" $O/b.err | tail -n 2 | cut -c 1-200
