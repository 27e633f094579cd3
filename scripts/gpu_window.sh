#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_mlp_epilogue.py -q 2>&1 | grep -E "^E  |passed|failed" | head -12 > gpurun_out/epilogue_tests.txt
cat gpurun_out/epilogue_tests.txt
