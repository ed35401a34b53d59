#!/bin/bash
# round 2, call 4E: per-kernel times of the three-kernel seeding stage (ncu launch list) + source counters of seed_walk_kernel
mkdir -p gpurun_out/r4e
O=gpurun_out/r4e
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum --clock-control none --kernel-name-base demangled --kernel-name regex:'pack_reads_kernel|seed_walk_kernel|plan_kernel' -c 12 --csv --log-file $O/seed_launches.csv python tools/experiments/stage_ab.py 4 ncu > $O/ncu1.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'seed_walk_kernel' -s 1 -c 1 -o $O/walk -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu2.log 2>&1; echo "ncu walk rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'plan_kernel' -s 1 -c 1 -o $O/plan -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu3.log 2>&1; echo "ncu plan rc=$?"
