#!/bin/bash
# round 2, call 6B: experiment -- two half-batches in flight on one GPU (two contexts, two host threads) against one whole batch
mkdir -p gpurun_out/r6b
timeout 300 python tools/experiments/two_lanes.py 4 > gpurun_out/r6b/out.txt 2> gpurun_out/r6b/err.txt; echo rc=$?
cat gpurun_out/r6b/out.txt; tail -n 3 gpurun_out/r6b/err.txt
