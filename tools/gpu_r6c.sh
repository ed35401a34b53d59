#!/bin/bash
# round 2, call 6C: what base alignment quality costs per 2 M pairs (option, outside the timed path)
mkdir -p gpurun_out/r6c
QM_AB_BAQ=3 timeout 600 python tools/experiments/stage_ab.py 4 "BAQ on TA-1-1" 2> gpurun_out/r6c/err.txt | tee gpurun_out/r6c/out.txt
timeout 300 python tools/experiments/stage_ab.py 4 "BAQ off TA-1-1" 2>> gpurun_out/r6c/err.txt | tee -a gpurun_out/r6c/out.txt
tail -n 3 gpurun_out/r6c/err.txt
