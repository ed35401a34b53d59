#!/bin/bash
# round 2, call J: where to hand over from rounds to the tail kernel
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
for t in 4096 8192 16384 32768 65536; do
  QM_TAIL_MIN=$t timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/tail_$t.json 2> $O/tail_$t.err; echo "tail $t rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j/tail_*.json'), key=lambda x:int(x.split('_')[-1].split('.')[0])):
    d=json.load(open(f)); print(f, round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stages_ms_per_step'].items() if v>0})
PY
