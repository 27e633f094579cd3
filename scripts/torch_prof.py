"""CUPTI (torch.profiler) kernel timings of graph-replayed minibatches: warm caches, real overlap."""
import os, sys, json
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from torchrl_b200.distributed import DataParallelContext
from torchrl_b200.networks import fused

mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
fused.set_matmul_mode(mode)
ctx = DataParallelContext()
class A:
    envs_per_gpu = bench.N_ENVS_PER_GPU
    no_graph = "--eager" in sys.argv
agent, col, buf, env = bench.build_agent(A, ctx, ctx.device)
for e in range(3):
    agent.current_epoch = e
    col.train_one_epoch(); agent.update_per_epoch()
torch.cuda.synchronize()
NMB = 8
agent._mb_state["upd"].zero_()
with fused.presplit(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NMB):
        agent._run_minibatch()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    k = e.name[:80]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
t0 = min(e.time_range.start for e in ev); t1 = max(e.time_range.end for e in ev)
print("mode", mode, "kernels", len(ev), "sum kernel us/mb", tot / NMB, "span us/mb", (t1 - t0) / NMB)
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-82s n/mb=%5.1f  us/mb=%8.1f  avg=%6.1f" % (k, c / NMB, t / NMB, t / c))
