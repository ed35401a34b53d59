#!/bin/bash
# round 2, call D: extend3 with sorted task lists + one-trip-ahead fetch; tests; bench on/off; per-launch metrics
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg2_ext3.json 2> $O/bench_cfg2_ext3.err; echo "ext3 rc=$?"
QM_EXT3=0 timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg2_ext2.json 2> $O/bench_cfg2_ext2.err; echo "ext2 rc=$?"
timeout 300 python bench.py --config 5 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > $O/bench_cfg5_ext3.json 2> $O/bench_cfg5_ext3.err; echo "cfg5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_cfg2.log 2>&1
ls -la $O
