#!/bin/bash
# round 2, call I: depth-cap tests + full suite, then compute-sanitizer (memcheck, racecheck) over the smoke run of the whole path
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -6 $O/pytest.log
timeout 1200 compute-sanitizer --tool memcheck --log-file $O/memcheck.log python -c "import __graft_entry__ as g; g.smoke()" > $O/memcheck.out 2>&1; echo "memcheck rc=$?"
tail -4 $O/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --log-file $O/racecheck.log python -c "import __graft_entry__ as g; g.smoke()" > $O/racecheck.out 2>&1; echo "racecheck rc=$?"
tail -4 $O/racecheck.log
ls -la $O
