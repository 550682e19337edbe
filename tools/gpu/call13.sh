#!/bin/bash
set -u
O=gpurun_out/r2m; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_cv.py -q -x > $O/pytest_sweep.log 2>&1; echo "rc=$?" >> $O/pytest_sweep.log
( time python bench.py --steps 10 --warmup 3 ) > $O/bench.json 2> $O/bench.err
tail -n 8 $O/*.log $O/*.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench.json').read().strip().splitlines()[0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','parity')}, 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'])
for k in ('evolved_population_value','ga_generation_s','evolution_c4_s','mc_c3_ms'): print(k, {a:b for a,b in d[k].items() if a not in ('what','note','workload')})
print({k:(v if not isinstance(v,dict) else v.get('value')) for k,v in d['cpu_baseline'].items() if k not in ('sample','configs0_backtest')})
PY
