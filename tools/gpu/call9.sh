#!/bin/bash
set -u
O=gpurun_out/r2i; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_cv.py tests/test_gpu_ga.py -q -x > $O/pytest_sweep.log 2>&1; echo "rc=$?" >> $O/pytest_sweep.log
timeout 300 python tools/tile_tune.py 0 8192 row_cost > $O/tpc.log 2>&1
B200BT_WARP_METRICS=1 timeout 300 python tools/tile_tune.py 0 8192 row_cost > $O/warpm.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -k "c2 or c5" > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
B200BT_ALLOW_SHORT_WARMUP=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 600 python tools/evolution_c4.py --generations 20 > $O/c4_20gen.json 2> $O/c4.err
tail -n 12 $O/*.log $O/*.err; cut -c1-600 $O/c4_20gen.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i/bench.json').read().strip().splitlines()[0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','parity')}, 'e2e', d['e2e']['value'])
for k in ('evolved_population_value','ga_generation_s','mc_c3_ms'): print(k, {a:b for a,b in d[k].items() if a not in ('what','note')})
PY
