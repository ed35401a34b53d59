#!/bin/bash
# round 2, call 6J: pieces of a chunk's bases in the host entry (8 chosen with the one-read-per-thread seeding kernel): e2e with 2 / 4 / 8
mkdir -p gpurun_out/r6j
O=gpurun_out/r6j
for p in 8 4 2; do
  QM_COPY_PARTS=$p timeout 300 python bench.py --cpu-seconds 0 --steps 10 --warmup 3 > $O/bench_p$p.json 2> $O/bench_p$p.err; echo "parts=$p rc=$?"
done
python - <<'PY'
import json
for p in (8, 4, 2):
    s = open(f"gpurun_out/r6j/bench_p{p}.json").read(); d = json.loads(s[s.index("{"):])
    print("parts", p, "resident", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), round(d["e2e"]["ms_per_step"], 2), "records", round(d["e2e"]["with_records"]["value"] / 1e6, 2))
PY
