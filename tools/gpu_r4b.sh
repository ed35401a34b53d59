#!/bin/bash
# round 2, call 4B: seeding kernel A/B -- filter batch 1 / 4 / 8, rows staged by one bulk copy or not; parity of the new build
mkdir -p gpurun_out/r4b
O=gpurun_out/r4b
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
for cfg in "1 0" "1 1" "4 1" "8 0" "8 1"; do
  set -- $cfg
  QM_BLOOM_BATCH=$1 QM_SEED_STAGE=$2 timeout 300 python tools/experiments/stage_ab.py 4 "batch=$1 stage=$2" 2>> $O/err.txt | tee -a $O/out.txt
done
QM_SEED_THREADS=128 timeout 300 python tools/experiments/stage_ab.py 4 "threads=128" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_THREADS=32 timeout 300 python tools/experiments/stage_ab.py 4 "threads=32" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 1 "default TA-50-1" 2>> $O/err.txt | tee -a $O/out.txt
tail -n 5 $O/err.txt
