#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_q.txt
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_r1_final.txt
