#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_checkpoint.py tests/test_offpolicy.py tests/test_sample_gather.py -q -m gpu 2>&1 | tail -60 > gpurun_out/new_tests.txt
tail -5 gpurun_out/new_tests.txt
