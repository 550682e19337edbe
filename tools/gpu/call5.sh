#!/bin/bash
set -u
O=gpurun_out/r2e; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_sweep.py -q -x -k "tiled or edge" > $O/pytest_tiled.log 2>&1; echo "rc=$?" >> $O/pytest_tiled.log
timeout 300 python tools/tile_tune.py 0,20,26,32 4096,8192 row_cost,row_thresholds > $O/tma.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -k "c2 or c5" > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py sweep ) > $O/racecheck.log 2>&1; echo "rc=$?" >> $O/racecheck.log
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py sweep ) > $O/memcheck.log 2>&1; echo "rc=$?" >> $O/memcheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane -f python tools/tile_profile.py 26 8192 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/lane.ncu-rep > $O/lane_summary.txt 2>&1
ncu -i $O/lane.ncu-rep --page source --csv > $O/lane_source.csv 2>/dev/null
ncu -i $O/lane.ncu-rep --page raw --csv > $O/lane_raw.csv 2>/dev/null
tail -n 30 $O/*.log; head -30 $O/lane_summary.txt
