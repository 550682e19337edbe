#!/bin/bash
set -u
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python tools/tile_tune.py 0,20,32,40 8192 row_cost > $O/tune.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sweep.py -q -x > $O/pytest_sweep.log 2>&1; echo "rc=$?" >> $O/pytest_sweep.log
python tools/evolved_profile.py 2>&1 | grep "ms per sweep" | cut -c1-70 > $O/evolved.log
tail -n 8 $O/*.log
