#!/bin/bash
set -u
O=gpurun_out/r2y; mkdir -p $O
B200BT_LS_CTAS=3 B200BT_LIB=gpurun_variants/ls3.so python tools/tile_tune.py 26,28,30,32,34 1536,2048,3072 > $O/tune_ls3.log 2>&1
for v in ls3s3 ls3u1 ls3u4; do
B200BT_LS_CTAS=3 B200BT_LIB=gpurun_variants/$v.so python tools/tile_tune.py 28,30,32 2048 > $O/tune_$v.log 2>&1
done
B200BT_LS_CTAS=2 B200BT_LIB=gpurun_variants/ls2.so python tools/tile_tune.py 20,24,30 2048 > $O/tune_ls2.log 2>&1
tail -n 17 $O/*.log
