#!/bin/bash
# round 2, call 4J: seed walk with one L2 round trip per trip (seed_walk_step2) -- parity, A/B, per-read histograms; new full-size tests
mkdir -p gpurun_out/r4j
O=gpurun_out/r4j
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py tests/test_fm_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
QM_SEED_ONETRIP=0 timeout 300 python tools/experiments/stage_ab.py 4 "onetrip=0 TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_DEBUG=1 timeout 300 python tools/experiments/stage_ab.py 4 "onetrip TA-1-1" 2> $O/err4.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 9 "onetrip TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 1 "onetrip TA-50-1" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_MINB=12 timeout 300 python tools/experiments/stage_ab.py 4 "onetrip minb=12 TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_REFILL=2 timeout 300 python tools/experiments/stage_ab.py 4 "onetrip refill=2 TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
grep "qm seed walk" $O/err4.txt | tail -n 13
timeout 1200 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q > $O/pytest_full.log 2>&1; echo "pytest full rc=$?"; tail -n 12 $O/pytest_full.log
