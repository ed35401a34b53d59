#!/bin/bash
# round 2, call C: one full ncu capture (with source counters) of ext3_kernel<80,32,true> on a round-0 launch of cfg2
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'ext3_kernel<.int.80,' -s 4 -c 1 -o $O/ext3_80 -f python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu.log 2>&1
ncu -i $O/ext3_80.ncu-rep --page source --csv --print-source sass > $O/ext3_80_sass.csv 2> $O/src.err
ncu -i $O/ext3_80.ncu-rep --page raw --csv > $O/ext3_80_raw.csv 2>> $O/src.err
ls -la $O
