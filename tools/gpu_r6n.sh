#!/bin/bash
# round 2, call 6N: ncu --set full with source of one baq_fast_kernel launch
mkdir -p gpurun_out/r6n
QM_AB_BAQ=3 timeout 200 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'baq_fast_kernel' -s 2 -c 1 -o gpurun_out/r6n/baq -f python tools/experiments/stage_ab.py 4 ncu > gpurun_out/r6n/ncu.log 2>&1; echo "rc=$?"
