#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_full.txt 2>&1
grep -v "^$" gpurun_out/dist_check_full.txt | grep -B2 -A25 "Traceback" | head -80
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --skip-roofline > gpurun_out/bench_n2_full.txt 2>&1
grep -B2 -A25 "Traceback" gpurun_out/bench_n2_full.txt | head -60; tail -2 gpurun_out/bench_n2_full.txt
