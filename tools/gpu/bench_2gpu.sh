#!/bin/bash
set -u
O=gpurun_out/bench_2gpu; mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench2.json 2> $O/bench2.err; echo "rc=$?" >> $O/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/ref2.json 2> $O/ref2.err; echo "rc=$?" >> $O/ref2.err
tail -n 12 $O/bench2.err | cut -c1-200; tail -n 3 $O/ref2.err | cut -c1-200; wc -c $O/*.json
