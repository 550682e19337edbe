#!/bin/bash
set -u
O=gpurun_out/r2u; mkdir -p $O
( python -m pytest tests/test_gpu_sweep.py tests/test_gpu_indicators.py -m gpu -q -x ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
python tools/indicator_bench.py --reps 20 2>&1 | head -3 > $O/ind_main.log
B200BT_LIB=gpurun_variants/rsi2.so python tools/indicator_bench.py --reps 20 2>&1 | head -3 > $O/ind_rsi2.log
python tools/e2e_breakdown.py > $O/e2e_main.log 2>&1
B200BT_LIB=gpurun_variants/rsi2.so python tools/e2e_breakdown.py > $O/e2e_rsi2.log 2>&1
python tools/tile_tune.py 30,32 8192 > $O/tune_main.log 2>&1
B200BT_LS_CTAS=3 B200BT_LIB=gpurun_variants/ls3.so python tools/tile_tune.py 24,30,32 8192 > $O/tune_ls3.log 2>&1
python bench.py --steps 10 --warmup 3 --skip-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -n 12 $O/*.log; python -c "
import json;d=json.load(open('$O/bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline'].get('dominant_kernel_ms'),d['roofline'].get('frac'),d['roofline'].get('dominant_kernel_share_of_step'))"
