#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_sample_gather.py -q -m gpu -k global_vec 2>&1 | tail -2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2>&1 | grep "dist_check ok\|Error" | tee gpurun_out/dist_check.txt
bash scripts/gpu_scale.sh 2
