#!/bin/bash
# round 2, call 4H: ncu --set full with source of the class 65-80 packed extension kernel (round 0 launch) at HEAD
mkdir -p gpurun_out/r4h
O=gpurun_out/r4h
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'ext3_kernel<\(int\)80' -s 2 -c 1 -o $O/ext3_80 -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 $O/ncu.log
