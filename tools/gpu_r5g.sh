#!/bin/bash
# round 2, call 5G: driver with flat per-file batches (no allocation per record): driver tests + file-level throughput
mkdir -p gpurun_out/r5g
O=gpurun_out/r5g
timeout 900 python -m pytest tests/test_driver_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python tools/file_level_bench.py 1000000 4 16 > $O/file_level_1m.jsonl 2> $O/file_level_1m.err; echo "file-level rc=$?"
cut -c1-330 $O/file_level_1m.jsonl
timeout 900 python tools/file_level_bench.py 4000000 4 16 > $O/file_level_4m.jsonl 2> $O/file_level_4m.err; echo "file-level rc=$?"
cut -c1-330 $O/file_level_4m.jsonl
