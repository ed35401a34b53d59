#!/bin/bash
# round 2, call U: pileup with wide loads; tests, bench, ncu --set full of pileup_kernel
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench",):
    s = open(f"gpurun_out/r2u/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, d["value"], d["ms_per_step"], d["e2e"] and d["e2e"]["value"], d["roofline"]["frac"])
    print(f, {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'pileup_kernel' -s 1 -c 1 -o $O/pileup -f python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
