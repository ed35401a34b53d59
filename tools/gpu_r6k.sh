#!/bin/bash
# round 2, call 6K: ncu launch list of the last build (bench.py --steps 1 --warmup 1)
mkdir -p gpurun_out/r6k
O=gpurun_out/r6k
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py $O/launches_cfg2.csv | head -n 12
