#!/bin/bash
set -u
O=gpurun_out/r3d; mkdir -p $O
( python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -m gpu -q -x ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
python tools/tile_tune.py 28 1024,1536,2048 > $O/tune.log 2>&1
for v in m26 m25; do B200BT_LIB=gpurun_variants/$v.so python tools/tile_tune.py 28 1536 > $O/tune_$v.log 2>&1; done
python tools/evolved_profile.py > $O/evolved.log 2>&1
for v in m26 m25; do B200BT_LIB=gpurun_variants/$v.so python tools/evolved_profile.py > $O/evolved_$v.log 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ll.csv python tools/tile_profile.py 0 1536 > $O/tp.log 2>&1
python - <<'PY'
import csv,re
rows=list(csv.reader(open('gpurun_out/r3d/ll.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
seq=[(re.sub(r'\(.*','',r[ki]).replace('b200bt::','').replace('void ',''), float(r[vi].replace(',',''))/1e3) for r in rows[hdr+2:] if len(r)>vi]
idx=[i for i,(k,v) in enumerate(seq) if k.startswith('lane_scan')]
for k,v in seq[idx[-1]-1:]: print(f"  {k:40s} {v:9.1f} us")
PY
tail -n 5 $O/*.log
