#!/bin/bash
set -u
O=gpurun_out/r3a; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/evolved_ll.csv -k regex:'lane_scan|chunk_|lane_combine|sweep_kernel' python tools/evolved_profile.py > $O/evolved.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:analyzer_c -c 1 -o $O/anc -f python tools/analyzer_bench.py > $O/anc.log 2>&1
python tools/ncu_summary.py $O/anc.ncu-rep > $O/anc_summary.txt 2>&1
ncu -i $O/anc.ncu-rep --page source --csv > $O/anc_source.csv 2>/dev/null
rm -f $O/anc.ncu-rep
python - <<'PY'
import csv,re,collections
rows=list(csv.reader(open('gpurun_out/r3a/evolved_ll.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(re.sub(r'\(.*','',r[ki]).replace('b200bt::','').replace('void ',''), float(r[vi].replace(',',''))/1e3) for r in rows[hdr+2:] if len(r)>vi]
idx=[i for i,(k,v) in enumerate(seq) if k.startswith('lane_scan')]
i0=idx[-1]
for k,v in seq[i0:i0+12]: print(f"  {k:40s} {v:9.1f} us")
c=collections.OrderedDict()
for k,v in seq[i0+12:]: c[k]=c.get(k,0)+v
for k,v in c.items(): print(f"  rest {k:35s} {v:9.1f} us")
PY
tail -3 $O/evolved.log; head -40 $O/anc_summary.txt
