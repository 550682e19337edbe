#!/bin/bash
set -u
O=gpurun_out/r3o; mkdir -p $O
( time timeout 240 python -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
( timeout 100 python __graft_entry__.py --smoke ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -n 8 $O/pytest.log; tail -n 3 $O/smoke.log
