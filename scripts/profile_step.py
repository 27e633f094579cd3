"""Profiling harness: warm up un-profiled, then run a short window of the PPO step between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py
The window holds: 8 collector steps, the epoch's GAE + old-log-prob pass, and 4 minibatch updates.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph", action="store_true", help="profile the graph-replay path (default eager)")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--minibatches", type=int, default=4)
    a = ap.parse_args()
    from torchrl_b200.distributed import DataParallelContext
    ctx = DataParallelContext()

    class A:
        envs_per_gpu = bench.N_ENVS_PER_GPU
        no_graph = not a.graph
    agent, col, buf, env = bench.build_agent(A, ctx, ctx.device)
    for epoch in range(2):
        agent.current_epoch = epoch
        col.train_one_epoch()
        agent.update_per_epoch()
    torch.cuda.synchronize()
    from torchrl_b200.networks import fused
    with fused.presplit():                      # what the collector / agent epochs run under (pre-split weight planes)
        torch.cuda.profiler.start()
        for _ in range(a.steps):
            col._step()
        agent.process_epoch_samples()
        agent._cache_old_logp()
        agent._mb_state["upd"].zero_()
        for _ in range(a.minibatches):
            agent._run_minibatch()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
