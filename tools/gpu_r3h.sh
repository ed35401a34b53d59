#!/bin/bash
# round 2, call 3H: the packed pieces are expanded on the seeding stream
mkdir -p gpurun_out/r3h
O=gpurun_out/r3h
timeout 600 python -m pytest tests/test_sample_gpu.py tests/test_driver_gpu.py tests/test_fm_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python tools/experiments/host_entry_stages.py > $O/out.txt 2> $O/err.txt; echo rc=$?
cat $O/out.txt
timeout 600 python bench.py --cpu-seconds 0 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config 4 --cpu-seconds 0 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_cfg4"):
    s = open(f"gpurun_out/r3h/{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4))
PY
