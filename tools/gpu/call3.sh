#!/bin/bash
# GPU call 3: full ncu profile (stall reasons, source hot spots) of lane_scan_kernel, warp-uniform zone skip, cost_row order
set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane -f python tools/tile_profile.py 26 8192 > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/lane.ncu-rep > $O/lane_summary.txt 2>&1
ncu -i $O/lane.ncu-rep --page raw --csv > $O/lane_raw.csv 2>/dev/null
ncu -i $O/lane.ncu-rep --page source --csv > $O/lane_source.csv 2>/dev/null
ls -la $O
