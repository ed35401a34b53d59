#!/bin/bash
# round 2, call 6F: filter batch size in the persistent walk (8 was chosen on the one-read-per-thread kernel)
mkdir -p gpurun_out/r6f
O=gpurun_out/r6f
for b in 2 4 6 8; do
  QM_BLOOM_BATCH=$b timeout 300 python tools/experiments/stage_ab.py 4 "batch=$b TA-1-1" 2>> $O/err.txt | tee -a $O/out.txt
done
QM_BLOOM_BATCH=4 timeout 300 python tools/experiments/stage_ab.py 1 "batch=4 TA-50-1" 2>> $O/err.txt | tee -a $O/out.txt
timeout 300 python tools/experiments/stage_ab.py 1 "batch=8 TA-50-1" 2>> $O/err.txt | tee -a $O/out.txt
