#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_final.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_final.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_final.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_final.txt
timeout 300 python scripts/torch_prof.py tc3 2>&1 | grep -v Warn | grep -A34 "^mode" > gpurun_out/torch_prof_final.txt
timeout 300 python scripts/phase_times.py 2>&1 | grep "^{" | tee gpurun_out/phase_times_final.txt
