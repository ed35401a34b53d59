#!/bin/bash
# round 2, call X: launch list with the speculative finish
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu.log 2>&1; echo "ncu rc=$?"
