#!/bin/bash
set -u
O=gpurun_out/bench_8gpu; mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench8.json 2> $O/bench8.err; echo "rc=$?" >> $O/bench8.err
grep -v "^\[bench" $O/bench8.err | tail -n 8 | cut -c1-200; wc -c $O/bench8.json
