#!/bin/bash
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-roofline --matmul tf32x3 2>&1 | tail -1 | tee gpurun_out/bench_n1_tf32x3.txt
