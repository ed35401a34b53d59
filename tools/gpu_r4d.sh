#!/bin/bash
# round 2, call 4D: seeding as pack / persistent walk / plan -- parity, A/B against the one-read-per-thread kernel, refill threshold
mkdir -p gpurun_out/r4d
O=gpurun_out/r4d
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
QM_SEED_PERSIST=0 timeout 300 python tools/experiments/stage_ab.py 4 "persist=0" 2>> $O/err.txt | tee -a $O/out.txt
for rf in 1 4 8 16; do
  QM_SEED_REFILL=$rf timeout 300 python tools/experiments/stage_ab.py 4 "persist refill=$rf" 2>> $O/err.txt | tee -a $O/out.txt
done
QM_SEED_THREADS=128 timeout 300 python tools/experiments/stage_ab.py 4 "persist threads=128" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_THREADS=32 timeout 300 python tools/experiments/stage_ab.py 4 "persist threads=32" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 1 "persist TA-50-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 9 "persist TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
tail -n 5 $O/err.txt
