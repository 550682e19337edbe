#!/bin/bash
set -u
O=gpurun_out/r2p; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane -f python tools/tile_profile.py 0 8192 > $O/ncu_lane.log 2>&1
python tools/ncu_summary.py $O/lane.ncu-rep > $O/lane_summary.txt 2>&1
ncu -i $O/lane.ncu-rep --page raw --csv > $O/lane_raw.csv 2>/dev/null
rm -f $O/lane.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:'chunk_partial|chunk_sums' -s 2 -c 2 -o $O/metrics -f python tools/tile_profile.py 0 8192 > $O/ncu_metrics.log 2>&1
python tools/ncu_summary.py $O/metrics.ncu-rep > $O/metrics_summary.txt 2>&1
rm -f $O/metrics.ncu-rep
B200BT_ALLOW_SHORT_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python tools/indicator_bench.py > $O/ind_bench.log 2>&1
head -24 $O/lane_summary.txt; grep -c lane_scan $O/launches.csv
