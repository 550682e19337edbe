#!/bin/bash
set -u
O=gpurun_out/r2n; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chunk_partial -s 4 -c 1 -o $O/partial -f python tools/evolved_profile.py > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/partial.ncu-rep > $O/partial_summary.txt 2>&1
ncu -i $O/partial.ncu-rep --page source --csv > $O/partial_source.csv 2>/dev/null
rm -f $O/partial.ncu-rep
head -40 $O/partial_summary.txt
