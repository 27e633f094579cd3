"""CUDA-event breakdown of one PPO epoch (graph path): rollout / prologue / update, per matmul mode."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torchrl_b200.distributed import DataParallelContext
from torchrl_b200.networks import fused

def run(mode, graph=True):
    fused.set_matmul_mode(mode)
    ctx = DataParallelContext()
    class A:
        envs_per_gpu = bench.N_ENVS_PER_GPU
        no_graph = not graph
    agent, col, buf, env = bench.build_agent(A, ctx, ctx.device)
    for e in range(3):
        agent.current_epoch = e
        col.train_one_epoch(); agent.update_per_epoch()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    tot = [0.0] * 4
    for it in range(3):
        ev[0].record(); col.rollout_no_sync()
        ev[1].record(); agent.process_epoch_samples(); agent._cache_old_logp()
        ev[2].record()
        st = agent._mb_state
        for _ in range(st["U"]):
            agent._run_minibatch()
        ev[3].record()
        torch.cuda.synchronize()
        for i in range(3):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
    print(json.dumps({"mode": mode, "graph": graph, "rollout_ms": tot[0] / 3, "prologue_ms": tot[1] / 3,
                      "update_ms": tot[2] / 3, "per_minibatch_us": tot[2] / 3 / st["U"] * 1e3,
                      "per_step_us": tot[0] / 3 / 128 * 1e3}), flush=True)

for mode in ("fp32", "tc3"):
    run(mode, True)
