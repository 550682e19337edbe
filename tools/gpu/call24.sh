#!/bin/bash
set -u
O=gpurun_out/r2x; mkdir -p $O
for w in 1024 2048; do
B200BT_LS_CTAS=3 B200BT_LIB=gpurun_variants/ls3.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ll_ls3_$w.csv python tools/tile_profile.py 30 $w > $O/tp_$w.log 2>&1
done
python - <<'PY'
import csv,re,glob
for f in sorted(glob.glob('gpurun_out/r2x/ll_*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
    seq=[(re.sub(r'\(.*','',r[ki]).replace('b200bt::','').replace('void ',''), float(r[vi].replace(',',''))/1e3) for r in rows[hdr+2:] if len(r)>vi]
    idx=[i for i,(k,v) in enumerate(seq) if k.startswith('lane_scan')]
    print(f, len(idx))
    i0=idx[-1]
    for k,v in seq[i0-1:i0+16]: print(f"  {k:40s} {v:9.1f} us")
    tail=seq[i0+16:]
    import collections
    c=collections.OrderedDict()
    for k,v in tail: c[k]=c.get(k,0)+v
    for k,v in c.items(): print(f"  rest {k:35s} {v:9.1f} us")
PY
