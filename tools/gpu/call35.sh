#!/bin/bash
set -u
O=gpurun_out/r3j; mkdir -p $O
( timeout 200 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -m gpu -q -x ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 420 python tools/hang_repro.py 90 25 > $O/plain.log 2>&1; echo "rc=$?" >> $O/plain.log
( timeout 200 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py sweep ) > $O/racecheck.log 2>&1; echo "rc=$?" >> $O/racecheck.log
python tools/tile_tune.py 28 1536 > $O/tune.log 2>&1
tail -n 3 $O/pytest.log $O/racecheck.log $O/tune.log; grep -c "gen" $O/plain.log; grep -v "stalls 0" $O/plain.log | tail -n 8 | cut -c1-200
