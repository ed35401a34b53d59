#!/bin/bash
# round 2, call 3A: ncu --set full of pair_decide_kernel, advance_kernel (round 0), cig_trace_kernel<128,128>
mkdir -p gpurun_out/r3a
O=gpurun_out/r3a
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'pair_decide_kernel|advance_kernel|cig_trace_kernel<.int.128' -s 3 -c 3 -o $O/three -f python bench.py --steps 1 --warmup 1 --no-e2e --cpu-seconds 0 > $O/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
