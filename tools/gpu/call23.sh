#!/bin/bash
set -u
O=gpurun_out/r2w; mkdir -p $O
( python -m pytest tests/test_gpu_sweep.py tests/test_gpu_fullsize.py -m gpu -q -x ) > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
python tools/tile_tune.py 30,36,44 1024,2048,4096 > $O/tune_main.log 2>&1
B200BT_LS_CTAS=3 B200BT_LIB=gpurun_variants/ls3.so python tools/tile_tune.py 24,30,36 2048,4096 > $O/tune_ls3.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py sweep ) > $O/memcheck.log 2>&1; echo "rc=$?" >> $O/memcheck.log
tail -n 16 $O/*.log
