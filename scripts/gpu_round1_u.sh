#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_u.txt
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1_tc3.txt
