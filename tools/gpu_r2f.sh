#!/bin/bash
# round 2, call F: full GPU test suite (indel table, driver outputs), default bench line, other configs
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "default rc=$?"
for c in 1 3 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"
done
timeout 900 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
ls -la $O
