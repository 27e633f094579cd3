#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_c.txt
python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
python bench.py --steps 3 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_n1.txt
python bench.py --steps 3 --warmup 3 --no-graph --skip-cpu-baseline --skip-roofline 2>&1 | tail -2 | tee gpurun_out/bench_n1_nograph.txt
