#!/bin/bash
set -u
O=gpurun_out/r2o; mkdir -p $O
python tools/evolved_profile.py 2>&1 | grep "ms per sweep" | cut -c1-60 > $O/base.log
for v in m6 m5 m4; do B200BT_LIB=$PWD/gpurun_variants/$v.so python tools/evolved_profile.py 2>&1 | grep "ms per sweep" | cut -c1-60 > $O/$v.log; done
tail -n 2 $O/*.log
