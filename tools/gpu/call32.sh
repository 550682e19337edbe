#!/bin/bash
set -u
O=gpurun_out/r3g; mkdir -p $O
( timeout 300 python -m pytest tests -m gpu -q -x ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( timeout 120 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
for i in 1 2 3 4 5; do
  B200BT_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --steps 10 --warmup 3 > $O/bench_$i.json 2> $O/bench_$i.err; echo "rc=$?" >> $O/bench_$i.err
done
B200BT_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --impl reference --steps 3 --warmup 1 > $O/ref.json 2> $O/ref.err; echo "rc=$?" >> $O/ref.err
tail -n 3 $O/pytest.log $O/smoke.log; tail -n 4 $O/bench_*.err $O/ref.err; wc -c $O/*.json
