#!/bin/bash
# round 2, call 4K: ncu --set full with source of the one-round-trip seed walk kernel
mkdir -p gpurun_out/r4k
O=gpurun_out/r4k
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'seed_walk_kernel' -s 1 -c 1 -o $O/walk2 -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu.log 2>&1; echo "ncu rc=$?"
