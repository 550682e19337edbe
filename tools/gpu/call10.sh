#!/bin/bash
set -u
O=gpurun_out/r2j; mkdir -p $O
timeout 300 python tools/tile_tune.py 0 8192 row_cost > $O/tune.log 2>&1
timeout 300 python -m pytest tests/test_gpu_sweep.py -q -x > $O/pytest_sweep.log 2>&1; echo "rc=$?" >> $O/pytest_sweep.log
B200BT_ALLOW_SHORT_WARMUP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
timeout 900 python tools/evolution_c4.py --generations 100 > $O/c4_100gen.json 2> $O/c4.err
timeout 900 python tools/evolution_c4.py --generations 100 --operators host > $O/c4_100gen_host.json 2>> $O/c4.err
tail -n 6 $O/*.log $O/*.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j/bench.json').read().strip().splitlines()[0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'])
for k in ('evolved_population_value','ga_generation_s','mc_c3_ms'): print(k, {a:b for a,b in d[k].items() if a not in ('what','note')})
for f in ('c4_100gen','c4_100gen_host'):
    c=json.load(open(f'gpurun_out/r2j/{f}.json')); print(f, {k:v for k,v in c.items() if k not in ('best_individual','workload')})
PY
