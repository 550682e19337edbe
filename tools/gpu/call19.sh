#!/bin/bash
set -u
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_indicators.py tests/test_gpu_backtest.py -q -x > $O/pytest_ind.log 2>&1; echo "rc=$?" >> $O/pytest_ind.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -k indicator > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
python tools/analyzer_bench.py 10 50 > $O/analyzer_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"analyzer|nanfill" --csv --log-file $O/an_launches.csv python tools/analyzer_bench.py 10 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2r/an_launches.csv")))
hi=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hi]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
seq=[(r[kn].split("(")[0][-28:], float(r[mv].replace(",",""))) for r in rows[hi+1:] if len(r)>mv]
for n,v in seq[-5:]: print(n, v/1e3, "us")
PY
tail -n 6 $O/*.log
