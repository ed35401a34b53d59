#!/bin/bash
# round 2, call 3F: do copies beside the kernels slow the kernels down?
mkdir -p gpurun_out/r3f
timeout 300 python tools/experiments/copy_interference.py 4 > gpurun_out/r3f/cfg4.txt 2>&1; cat gpurun_out/r3f/cfg4.txt | tail -n 6
timeout 300 python tools/experiments/copy_interference.py 2 > gpurun_out/r3f/cfg2.txt 2>&1; cat gpurun_out/r3f/cfg2.txt | tail -n 6
