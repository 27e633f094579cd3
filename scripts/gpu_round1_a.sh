#!/bin/bash
# first GPU contact: device info, GAE parity, sanitizer, microbench, ncu
set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gae.py -x -q -m gpu -k "golden or empty" 2>&1 | tail -8 | tee gpurun_out/sanitizer.txt
python scripts/gae_bench.py 2>&1 | tee gpurun_out/gae_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_chunked -s 3 -c 2 -o gpurun_out/gae_prof_r1 python scripts/gae_bench.py --sizes 128x1048576 --iters 2 --variants 1 > gpurun_out/ncu_gae.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_chunked -s 3 -c 2 -o gpurun_out/gae_prof_small_r1 python scripts/gae_bench.py --sizes 128x4096 --iters 2 --variants 1 >> gpurun_out/ncu_gae.log 2>&1
ls -la gpurun_out
