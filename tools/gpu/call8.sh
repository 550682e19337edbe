#!/bin/bash
set -u
O=gpurun_out/r2h; mkdir -p $O
( time python -m pytest tests -m gpu -q --durations=8 ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( time python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
( time python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
for tool in memcheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_target.py ) > $O/sanitize_$tool.log 2>&1
  echo "rc=$?" >> $O/sanitize_$tool.log
done
B200BT_ALLOW_SHORT_WARMUP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-extras --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
tail -5 $O/pytest.log $O/smoke.log $O/bench.err $O/sanitize_*.log; cut -c1-1500 $O/bench.json
