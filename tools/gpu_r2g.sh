#!/bin/bash
# round 2, call G: GPU test suite with the FM-index seeder
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
