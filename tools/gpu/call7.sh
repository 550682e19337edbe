#!/bin/bash
set -u
O=gpurun_out/r2g; mkdir -p $O
timeout 300 python tools/tile_tune.py 0,20,32 8192 row_cost > $O/fine.log 2>&1
B200BT_LIB=$PWD/gpurun_variants/s3.so timeout 300 python tools/tile_tune.py 0 8192 row_cost > $O/fine_s3.log 2>&1
timeout 300 python -m pytest tests/test_gpu_sweep.py -q -x -k "tiled or edge" > $O/pytest_tiled.log 2>&1; echo "rc=$?" >> $O/pytest_tiled.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane -f python tools/tile_profile.py 26 8192 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/lane.ncu-rep > $O/lane_summary.txt 2>&1
ncu -i $O/lane.ncu-rep --page source --csv > $O/lane_source.csv 2>/dev/null
ncu -i $O/lane.ncu-rep --page raw --csv > $O/lane_raw.csv 2>/dev/null
tail -n 12 $O/*.log; head -26 $O/lane_summary.txt
