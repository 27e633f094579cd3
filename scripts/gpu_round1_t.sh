#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_mlp_epilogue.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_tc3.txt
timeout 300 python scripts/phase_times.py 2>&1 | grep "^{" | tee gpurun_out/phase_times_tc3.txt
