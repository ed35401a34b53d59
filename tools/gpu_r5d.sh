#!/bin/bash
# round 2, call 5D: small batches (config 1, 100 k pairs per step): thresholds of the tail and of the speculative finish relative to the batch
mkdir -p gpurun_out/r5d
O=gpurun_out/r5d
run() { echo "== $1"; env $2 timeout 300 python bench.py --config 1 --cpu-seconds 0 --no-e2e 2> $O/err.txt | python -c "
import sys, json
s = sys.stdin.read(); d = json.loads(s[s.index('{'):])
print(round(d['value'] / 1e6, 2), round(d['ms_per_step'], 2), {k: round(v, 2) for k, v in d['stages_ms_per_step'].items()})"; }
run "defaults" "QM_NOP=1"
run "tail 2048 spec 12500" "QM_TAIL_MIN=2048 QM_SPEC_MIN=12500"
run "tail 4096 spec 25000" "QM_TAIL_MIN=4096 QM_SPEC_MIN=25000"
run "tail 2048 spec off" "QM_TAIL_MIN=2048 QM_SPEC=0"
run "tail 8192 spec off" "QM_TAIL_MIN=8192 QM_SPEC=0"
run "tail 1024 spec 6000" "QM_TAIL_MIN=1024 QM_SPEC_MIN=6000"
