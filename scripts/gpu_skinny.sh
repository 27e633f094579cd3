#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/skinny_bench.py 16384 > gpurun_out/skinny_bench.txt 2>&1
timeout 400 python -m pytest tests/test_mlp_epilogue.py tests/test_ppo_pipeline.py tests/test_reference_parity.py tests/test_ppo_loss_optim.py -q -m gpu 2>&1 | tail -25 >> gpurun_out/skinny_bench.txt
timeout 400 python bench.py > gpurun_out/bench_skinny.txt 2>&1
tail -c 1500 gpurun_out/bench_skinny.txt >> gpurun_out/skinny_bench.txt
cat gpurun_out/skinny_bench.txt
