#!/bin/bash
set -u
O=gpurun_out/r2k; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep.py -q -x > $O/pytest_sweep.log 2>&1; echo "rc=$?" >> $O/pytest_sweep.log
B200BT_ALLOW_SHORT_WARMUP=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-extras > $O/bench.json 2> $O/bench.err
tail -n 25 $O/*.log $O/*.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k/bench.json').read().strip().splitlines()[0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], d['e2e']['value']/d['value'])
PY
