#!/bin/bash
# round 2, call 4F: BAQ parity; seeding walk without local-memory state, four-at-a-time seed scan, adaptive filter batch; register budgets
mkdir -p gpurun_out/r4f
O=gpurun_out/r4f
timeout 900 python -m pytest tests/test_baq_gpu.py -m gpu -x -q > $O/pytest_baq.log 2>&1; echo "pytest baq rc=$?"; tail -n 12 $O/pytest_baq.log
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
for mb in 8 12 16; do
  QM_SEED_MINB=$mb timeout 300 python tools/experiments/stage_ab.py 4 "minb=$mb" 2>> $O/err.txt | tee -a $O/out.txt
done
QM_SEED_MINB=12 QM_SEED_REFILL=2 timeout 300 python tools/experiments/stage_ab.py 4 "minb=12 refill=2" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_MINB=12 timeout 300 python tools/experiments/stage_ab.py 9 "minb=12 TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
tail -n 5 $O/err.txt
