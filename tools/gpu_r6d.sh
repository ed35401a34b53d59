#!/bin/bash
# round 2, call 6D: BAQ with the live HMM rows in shared memory -- parity and cost; seed_find with 8-byte loads
mkdir -p gpurun_out/r6d
O=gpurun_out/r6d
timeout 900 python -m pytest tests/test_baq_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 $O/pytest.log
timeout 600 python -m pytest tests/test_driver_gpu.py -m gpu -x -q -k baq > $O/pytest_drv.log 2>&1; echo "pytest driver baq rc=$?"; tail -n 2 $O/pytest_drv.log
QM_AB_BAQ=3 timeout 600 python tools/experiments/stage_ab.py 4 "BAQ on TA-1-1" 2> $O/err.txt | tee $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 4 "BAQ off TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
