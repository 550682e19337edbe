#!/bin/bash
set -u
O=gpurun_out/r2d; mkdir -p $O
python tools/tile_tune.py 0 8192 cost,cost_row > $O/wz.log 2>&1
B200BT_LIB=$PWD/gpurun_variants/fz2.so python tools/tile_tune.py 0 8192 cost,cost_row > $O/fz2.log 2>&1
B200BT_LIB=$PWD/gpurun_variants/fz2.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane_fz2 -f python tools/tile_profile.py 26 8192 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/lane_fz2.ncu-rep > $O/lane_fz2_summary.txt 2>&1
ncu -i $O/lane_fz2.ncu-rep --page raw --csv > $O/lane_fz2_raw.csv 2>/dev/null
tail -n +1 $O/*.log; head -30 $O/lane_fz2_summary.txt
