#!/bin/bash
# round 2, call 4I: seed walk with fewer dependent stages per trip (warp ranges, unique positions inline, no needless seed scans, sort
# in the plan kernel); extension: four-wide trimming scans, prefetch in the blended columns.  Parity + stage times.
mkdir -p gpurun_out/r4i
O=gpurun_out/r4i
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_sample_gpu.py tests/test_extend_gpu.py tests/test_fm_gpu.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
QM_SEED_DEBUG=1 timeout 300 python tools/experiments/stage_ab.py 4 "TA-1-1" 2> $O/err4.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 4 "TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 9 "TA-0-1" 2>> $O/err.txt | tee -a $O/out.txt
QM_SEED_REFILL=4 timeout 300 python tools/experiments/stage_ab.py 4 "TA-1-1 refill=4" 2>> $O/err.txt | tee -a $O/out.txt
grep "qm seed walk" $O/err4.txt | tail -n 13
