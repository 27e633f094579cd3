#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_tc.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/pytest_gemm_tc.txt
timeout 120 python scripts/gemm_tc_bench.py 2>&1 | tail -8 | tee gpurun_out/gemm_tc_bench.txt
