#!/bin/bash
set -u
O=gpurun_out/r3l; mkdir -p $O
( timeout 200 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -m gpu -q -x ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 100 python tools/tile_tune.py 28 1536 > $O/tune.log 2>&1
tail -n 6 $O/pytest.log $O/tune.log
