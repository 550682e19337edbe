#!/bin/bash
set -u
O=gpurun_out/stress; mkdir -p $O
timeout 330 python tools/hang_repro.py 70 25 > $O/plain.log 2>&1; echo "rc=$?" >> $O/plain.log
for i in 1 2; do B200BT_BENCH_WATCHDOG_S=200 timeout 260 python bench.py --steps 10 --warmup 3 > $O/bench_$i.json 2> $O/bench_$i.err; echo "rc=$?" >> $O/bench_$i.err; done
grep -c "gen" $O/plain.log; grep -v "stalls 0" $O/plain.log | tail -n 6 | cut -c1-200; tail -n 2 $O/bench_*.err; wc -c $O/*.json
