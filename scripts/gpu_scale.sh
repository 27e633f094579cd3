#!/bin/bash
# usage: gpu_scale.sh N
N=$1
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --skip-roofline > gpurun_out/bench_n${N}_full.txt 2>&1
echo "exit code $?" >> gpurun_out/bench_n${N}_full.txt
tail -2 gpurun_out/bench_n${N}_full.txt | cut -c1-300
