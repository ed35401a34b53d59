#!/bin/bash
# round 2, call 4G: where the seed walk's time goes, read by read (QM_SEED_DEBUG histograms), clean sample and 1:1 mixture
mkdir -p gpurun_out/r4g
O=gpurun_out/r4g
QM_SEED_DEBUG=1 timeout 300 python tools/experiments/stage_ab.py 9 "debug TA-0-1" 2> $O/err9.txt | tee -a $O/out.txt
QM_SEED_DEBUG=1 timeout 300 python tools/experiments/stage_ab.py 4 "debug TA-1-1" 2> $O/err4.txt | tee -a $O/out.txt
grep "qm seed walk" $O/err9.txt | tail -n 40
grep "qm seed walk" $O/err4.txt | tail -n 40
