#!/bin/bash
# Round-1 closing run on ONE B200: full GPU test suite, smoke, bench, profiles of the final build.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_final.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_final.txt
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_final.txt
# does nvidia-smi polling perturb the timed region?  4 extra pollers (what a 4-rank job used to start)
for i in 1 2 3 4; do
  timeout 40 nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 100 > /dev/null 2>&1 &
done
timeout 300 python bench.py --steps 3 --warmup 3 --skip-roofline --skip-cpu-baseline 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench_n1_pollers.txt
wait
timeout 300 python scripts/torch_prof.py tc3 2>&1 | grep -v Warn | grep -A40 "^mode" > gpurun_out/torch_prof_final.txt
timeout 300 python scripts/phase_times.py 2>&1 | grep "^{" | tee gpurun_out/phase_times_final.txt
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_ppo_final_r1.csv python scripts/profile_step.py > gpurun_out/ncu_list.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"skinny" -c 16 -o gpurun_out/skinny_kernels_r1 python scripts/profile_step.py --steps 1 --minibatches 1 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -12
