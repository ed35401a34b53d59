#!/bin/bash
# round 2, call 5F: driver host side in parallel (mate files read side by side, packing and BAM record encoding on -t threads, the next
# batch read while the current one is on the device): driver tests + file-level throughput, 1 M and 4 M pairs
mkdir -p gpurun_out/r5f
O=gpurun_out/r5f
timeout 900 python -m pytest tests/test_driver_gpu.py tests/test_dedup_gpu.py tests/test_depthcap_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python tools/file_level_bench.py 1000000 4 16 > $O/file_level_1m.jsonl 2> $O/file_level_1m.err; echo "file-level rc=$?"
cat $O/file_level_1m.jsonl
timeout 900 python tools/file_level_bench.py 4000000 4 16 > $O/file_level_4m.jsonl 2> $O/file_level_4m.err; echo "file-level rc=$?"
cat $O/file_level_4m.jsonl
nproc; free -g | head -n 2
