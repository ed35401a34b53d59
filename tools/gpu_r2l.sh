#!/bin/bash
# round 2, call L: packed host entry — GPU tests, the default bench, cfg 3 and cfg 1 e2e
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --config 1 --cpu-seconds 0 > $O/bench_cfg1.json 2> $O/bench_cfg1.err; echo "cfg1 rc=$?"
timeout 300 python bench.py --config 3 --cpu-seconds 0 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
head -c 1500 $O/bench_default.json; echo; tail -n 3 $O/bench_default.err
