#!/bin/bash
# round 2, call S: launch list of the current build (cfg 2)
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -n 2 $O/ncu.log | cut -c 1-300
