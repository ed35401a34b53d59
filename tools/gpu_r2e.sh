#!/bin/bash
# round 2, call E: extend3 byte planes vs 16-bit planes; tests; source-level capture
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg2_narrow.json 2> $O/bench_cfg2_narrow.err; echo "narrow rc=$?"
QM_EXT3_NARROW=0 timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg2_wide.json 2> $O/bench_cfg2_wide.err; echo "wide rc=$?"
QM_TPT_MIN=4096 timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_cfg2_narrow_min4k.json 2> $O/bench_cfg2_narrow_min4k.err; echo "min4k rc=$?"
timeout 300 python bench.py --config 5 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed_pipe_alu.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_cfg2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'ext3_kernel<.int.80,' -s 2 -c 1 -o $O/ext3_80 -f python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1
ncu -i $O/ext3_80.ncu-rep --page source --csv --print-source sass > $O/ext3_80_sass.csv 2> $O/src.err
ncu -i $O/ext3_80.ncu-rep --page raw --csv > $O/ext3_80_raw.csv 2>> $O/src.err
rm -f $O/ext3_80.ncu-rep
ls -la $O
