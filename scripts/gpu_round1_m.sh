#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_m.txt
python scripts/phase_times.py 2>&1 | grep "^{" | tee gpurun_out/phase_times.txt
python scripts/torch_prof.py fp32 2>&1 | grep -v Warn | tail -26 | tee gpurun_out/torch_prof_fp32.txt
