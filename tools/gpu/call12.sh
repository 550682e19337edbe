#!/bin/bash
set -u
O=gpurun_out/r2l; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lane_scan|chunk_|lane_combine|sweep_kernel' --csv --log-file $O/evolved_launches.csv python tools/evolved_profile.py > $O/evolved.log 2>&1
tail -3 $O/evolved.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2l/evolved_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
seq=[(r[kn].split('(')[0][-40:], float(r[mv].replace(',',''))) for r in rows[hi+1:] if len(r)>mv]
idx=[i for i,(n,v) in enumerate(seq) if 'lane_scan' in n]
start=idx[-1]
t=collections.OrderedDict(); c=collections.Counter()
for n,v in seq[start:]: t[n]=t.get(n,0)+v; c[n]+=1
for k,v in t.items(): print(f"{k:42s} x{c[k]:3d} {v/1e6:8.3f} ms")
print("total", sum(t.values())/1e6)
PY
