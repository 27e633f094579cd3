#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ppo_eager_r1c.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tf32x3|splitk_reduce|bias_act_bwd|synth_env_step" -c 12 -o gpurun_out/tc_kernels_r1 python scripts/profile_step.py --steps 2 --minibatches 1 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
