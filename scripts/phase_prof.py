"""Where one PPO step (BASELINE.json configs[1]) spends its time: the collector epoch, the per-epoch preparation
(GAE, old log-probs, advantage statistics, permutation upload) and the minibatch loop, each timed with CUDA events on
warm graphs; plus the CUPTI kernel table of graph-replayed COLLECTOR steps (the minibatch table is scripts/torch_prof.py).

    python scripts/phase_prof.py
"""
import os
import sys
import time

import torch
from torch.profiler import profile, ProfilerActivity

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from torchrl_b200.distributed import DataParallelContext  # noqa: E402
from torchrl_b200.networks import fused  # noqa: E402

ctx = DataParallelContext()
if ctx.rank != 0:
    sys.stdout = open(os.devnull, "w")


class A:
    envs_per_gpu = bench.N_ENVS_PER_GPU
    no_graph = False


agent, col, buf, env = bench.build_agent(A, ctx, ctx.device)
for e in range(3):
    agent.current_epoch = e
    col.train_one_epoch()
    agent.update_per_epoch()
torch.cuda.synchronize()


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3


for rep in range(3):
    agent.current_epoch = 3 + rep
    c_dev, c_wall = timed(col.train_one_epoch)
    u_dev, u_wall = timed(agent.update_per_epoch)
    print(f"epoch {rep}: collector {c_dev:.2f} ms (wall {c_wall:.2f}) | update_per_epoch {u_dev:.2f} ms (wall {u_wall:.2f})")

# inside update_per_epoch: everything before the minibatch loop
if hasattr(agent, "_run_minibatch"):
    U = agent._mb_state["U"]

    def loop():
        agent._mb_state["upd"].zero_()
        with fused.presplit():
            for _ in range(U):
                agent._run_minibatch()
    l_dev, l_wall = timed(loop)
    print(f"minibatch loop alone: {U} x {l_dev / U * 1e3:.1f} us = {l_dev:.2f} ms (wall {l_wall:.2f})")

NST = 16
with fused.presplit(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NST):
        col._step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    a = agg.setdefault(e.name[:80], [0, 0.0])
    a[0] += 1
    a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in ev)
t1 = max(e.time_range.end for e in ev)
print("collector step: kernels/step", len(ev) / NST, "sum kernel us/step", tot / NST, "span us/step", (t1 - t0) / NST)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print("%-82s n/step=%5.1f  us/step=%8.1f  avg=%6.1f" % (k, c / NST, t / NST, t / c))
