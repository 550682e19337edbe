#!/bin/bash
# GPU call 1 (round 2): full-size parity, smoke, sanitizers, ncu evidence for families 1 and 3, baseline bench
set -u
O=gpurun_out/r2a; mkdir -p $O
nproc > $O/nproc.txt; nvidia-smi -L >> $O/nproc.txt
( time python -m pytest tests -m gpu -q --durations=15 ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( time python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
for tool in memcheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_target.py ) > $O/sanitize_$tool.log 2>&1
  echo "rc=$?" >> $O/sanitize_$tool.log
done
# ncu: families 3 and 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mc_paths|sel_|mc_moments' -c 12 -o $O/mc -f python tools/mc_bench.py --reps 1 > $O/ncu_mc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'bank_kernel|macd|bollinger|extrema|vwap|nanfill|sma' -c 80 -o $O/ind -f python tools/indicator_bench.py --reps 1 --warmup 0 > $O/ncu_ind.log 2>&1
for r in mc ind; do python tools/ncu_summary.py $O/$r.ncu-rep > $O/${r}_ncu_summary.txt 2>&1; done
rm -f $O/ind.ncu-rep; [ $(stat -c %s $O/mc.ncu-rep) -gt 30000000 ] && rm -f $O/mc.ncu-rep
python tools/mc_bench.py > $O/mc_bench.json 2>&1
python tools/indicator_bench.py > $O/ind_bench.log 2>&1
ls -la $O
