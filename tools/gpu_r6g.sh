#!/bin/bash
# round 2, call 6G: seeding filter sized by the genome (16 bits per k-mer in L2; config 3's 4.9 Mb index had none): parity, configs 2 and 3
mkdir -p gpurun_out/r6g
O=gpurun_out/r6g
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 300 python tools/experiments/stage_ab.py 4 "filter 16 bits/k-mer TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 9 "filter 16 bits/k-mer TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 600 python bench.py --config 3 --cpu-seconds 0 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json
s = open("gpurun_out/r6g/bench_cfg3.json").read(); d = json.loads(s[s.index("{"):])
print(round(d["value"] / 1e6, 2), round(d["ms_per_step"], 2), round(d["e2e"]["value"] / 1e6, 2), {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
