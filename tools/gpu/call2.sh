#!/bin/bash
# GPU call 2: dispatch-order / warp-zone / fine-zone / K / warm experiments on the C2 workload
set -u
O=gpurun_out/r2b; mkdir -p $O
B200BT_LIB=$PWD/gpurun_variants/warpzone.so python tools/tile_tune.py 0 8192 cost,period,row,rowth,cost_row > $O/wz_orders.log 2>&1
python tools/tile_tune.py 0 8192 cost,period,row,rowth,cost_row > $O/fz_orders.log 2>&1
python tools/tile_tune.py 13,20,26,32 2048,4096,8192 cost,cost_row > $O/fz_kwarm.log 2>&1
B200BT_LIB=$PWD/gpurun_variants/fz_t128.so B200BT_LS_THREADS=128 B200BT_LS_CTAS=8 python tools/tile_tune.py 0,13,20,26 4096,8192 cost,cost_row,rowth > $O/fz_t128.log 2>&1
python -m pytest tests/test_gpu_sweep.py -q -x -k "tiled or edge" > $O/pytest_tiled.log 2>&1
tail -n +1 $O/*.log
