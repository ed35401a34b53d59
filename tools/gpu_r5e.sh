#!/bin/bash
# round 2, call 5E: file-level throughput of qm_driver sample (FASTQ -> every output of the replaced rules); config 1 with the new thresholds
mkdir -p gpurun_out/r5e
O=gpurun_out/r5e
timeout 600 python tools/file_level_bench.py 1000000 4 16 > $O/file_level.jsonl 2> $O/file_level.err; echo "file-level rc=$?"
cat $O/file_level.jsonl; tail -n 2 $O/file_level.err
timeout 300 python bench.py --config 1 --cpu-seconds 0 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "cfg1 rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r5e/bench_cfg1.json").read(); d = json.loads(s[s.index("{"):])
print(round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -n 2
