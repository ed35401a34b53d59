#!/bin/bash
# round 2, call A: GPU tests, bench lines of all five BASELINE configs, round log + launch list of cfg2 / cfg5
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for c in 2 1 3 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"
done
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
QM_ROUND_LOG=1 timeout 300 python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/roundlog_cfg2.json 2> $O/roundlog_cfg2.err
QM_ROUND_LOG=1 timeout 300 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/roundlog_cfg5.json 2> $O/roundlog_cfg5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg2.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_cfg2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_cfg5.csv python bench.py --config 5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_cfg5.log 2>&1
ls -la $O
