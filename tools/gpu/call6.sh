#!/bin/bash
set -u
O=gpurun_out/r2f; mkdir -p $O
timeout 300 python tools/tile_tune.py 0,20,26,32 4096,8192 row_cost,row_thresholds > $O/wp.log 2>&1
for v in u1 s2 s4; do B200BT_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/tile_tune.py 0 8192 row_cost > $O/wp_$v.log 2>&1; done
timeout 300 python -m pytest tests/test_gpu_sweep.py -q -x -k "tiled or edge" > $O/pytest_tiled.log 2>&1; echo "rc=$?" >> $O/pytest_tiled.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py -q -x -k "c2 or c5" > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
tail -n 30 $O/*.log
