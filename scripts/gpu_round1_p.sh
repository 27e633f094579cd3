#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_examples.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_examples.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ppo_eager_r1b.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"synth_env_step|collect_finalize|bias_act_bwd|bias_act_fwd|ppo_actor_loss|ppo_critic_loss|row_copy|adam_step|grad_sumsq|tanh_gaussian|obs_filt|vec_stats" -o gpurun_out/trl_kernels_r1 python scripts/profile_step.py --steps 2 --minibatches 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -8
