#!/bin/bash
# round 2, call 4C: ncu --set full of seed_chain_kernel on the TA-1-1 sample; DRAM traffic of pileup_kernel and seed_chain_kernel
mkdir -p gpurun_out/r4c
O=gpurun_out/r4c
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'seed_chain_kernel' -s 1 -c 1 -o $O/seed -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu_seed.log 2>&1; echo "ncu seed rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled --kernel-name regex:'pileup_kernel|seed_chain_kernel' -s 2 -c 4 --csv --log-file $O/traffic.csv python tools/experiments/stage_ab.py 4 ncu > $O/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
ls -la $O
