#!/bin/bash
# round 2, call 3G: stage timers, resident entry against host entry, one chunk and four
mkdir -p gpurun_out/r3g
QM_HOST_TRACE=1 timeout 600 python tools/experiments/host_entry_stages.py > gpurun_out/r3g/out.txt 2> gpurun_out/r3g/err.txt; echo rc=$?
cat gpurun_out/r3g/out.txt; grep "host trace" gpurun_out/r3g/err.txt | tail -n 3 | cut -c 1-300; tail -n 3 gpurun_out/r3g/err.txt | cut -c 1-300
