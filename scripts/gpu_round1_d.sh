#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/cpu_threads_probe.py 2>&1 | tee gpurun_out/cpu_threads.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ppo_eager_r1.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log
wc -l gpurun_out/launches_ppo_eager_r1.csv
