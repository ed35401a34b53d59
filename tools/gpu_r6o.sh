#!/bin/bash
# round 2, call 6O: BAQ with two live rows + compact decode matrix: parity and cost
mkdir -p gpurun_out/r6o
timeout 120 python -m pytest tests/test_baq_gpu.py -m gpu -x -q 2>&1 | tail -n 2
QM_AB_BAQ=3 timeout 120 python tools/experiments/stage_ab.py 4 "BAQ on, live rows + compact matrix" 2> gpurun_out/r6o/err.txt | tee gpurun_out/r6o/out.txt
