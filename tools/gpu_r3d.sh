#!/bin/bash
# round 2, call 3D: tail below 32768 tasks again; device-wide scan in qm_call_snps; cfg 4 host entry chunk by chunk
mkdir -p gpurun_out/r3d
O=gpurun_out/r3d
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/pytest.log
for c in 1 3; do
  timeout 600 python bench.py --config $c --cpu-seconds 0 > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"
done
QM_HOST_TRACE=1 timeout 600 python bench.py --config 4 --cpu-seconds 0 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
grep "host trace" $O/bench_cfg4.err | tail -n 4 | cut -c 1-400
python - <<'PY'
import json
for f in ("cfg1", "cfg3", "cfg4"):
    s = open(f"gpurun_out/r3d/bench_{f}.json").read(); d = json.loads(s[s.index("{"):])
    print(f, round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), d["e2e"] and round(d["e2e"]["value"] / 1e6, 2), round(d["roofline"]["frac"], 4), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
