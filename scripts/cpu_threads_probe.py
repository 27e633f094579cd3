"""How fast is the CPU port's PPO minibatch (16384 samples, MLP(256,256)) vs torch thread count?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ref_port
import torch.nn as nn
N = 4096
for threads in (8, 16, 32, 64, 128):
    torch.set_num_threads(threads)
    torch.manual_seed(0); np.random.seed(0)
    pf = ref_port.TanhGaussianPolicy(17, 6, [256, 256], nn.Tanh)
    vf = ref_port.MLPNet(17, 1, [256, 256], nn.Tanh)
    buf = ref_port.Float64Rollout(8 * N, N, True)
    agent = ref_port.PPOPort(pf, vf, buf, batch_size=4 * N)
    B = 4 * N
    batch = {"obs": np.random.randn(B, 17), "acts": np.tanh(np.random.randn(B, 6)), "advs": np.random.randn(B, 1),
             "estimate_returns": np.random.randn(B, 1), "values": np.random.randn(B, 1)}
    agent.update(batch)
    t0 = time.perf_counter()
    for _ in range(3):
        agent.update(batch)
    print("threads", threads, "minibatch_s", (time.perf_counter() - t0) / 3, flush=True)
