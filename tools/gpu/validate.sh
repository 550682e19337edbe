#!/bin/bash
set -u
O=gpurun_out/validate; mkdir -p $O
( time timeout 240 python -m pytest tests -m gpu -q --durations=6 ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( time timeout 120 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time timeout 300 python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:lane_scan -s 1 -c 1 -o $O/lane -f python tools/tile_profile.py > $O/ncu_lane.log 2>&1
python tools/ncu_summary.py $O/lane.ncu-rep > $O/lane_summary.txt 2>&1
ncu -i $O/lane.ncu-rep --page raw --csv > $O/lane_raw.csv 2>/dev/null
rm -f $O/lane.ncu-rep
B200BT_ALLOW_SHORT_WARMUP=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
for tool in memcheck racecheck; do
  ( time timeout 240 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_target.py ) > $O/sanitize_$tool.log 2>&1
  echo "rc=$?" >> $O/sanitize_$tool.log
done
timeout 120 python tools/indicator_bench.py --reps 20 > $O/indicator_bench.txt 2>&1
timeout 120 python tools/analyzer_bench.py > $O/analyzer_bench.txt 2>&1
timeout 120 ncu --set full --clock-control none -k regex:rsi_bank -c 1 -o $O/rsi -f python tools/indicator_bench.py --reps 1 --warmup 0 > $O/ncu_rsi.log 2>&1
python tools/ncu_summary.py $O/rsi.ncu-rep > $O/rsi_summary.txt 2>&1; rm -f $O/rsi.ncu-rep
tail -n 4 $O/pytest.log $O/smoke.log $O/sanitize_*.log; head -22 $O/lane_summary.txt; head -3 $O/indicator_bench.txt; cat $O/analyzer_bench.txt; tail -3 $O/bench.err
