#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --skip-roofline > gpurun_out/bench_n2_full.txt 2>&1
echo "exit code $?" >> gpurun_out/bench_n2_full.txt
tail -2 gpurun_out/bench_n2_full.txt | cut -c1-400
