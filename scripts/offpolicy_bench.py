"""Secondary measurements (not the bench.py headline): BASELINE configs 3 and 4 on one GPU.

  config 3: TwinSAC-Q, 1024 SynthAnt envs (obs 111, act 8), 1M-transition ring, batch 4096, MLP(256,256)
  config 4: QR-DQN (200 quantiles), 512 SynthAtari envs (4x84x84 uint8), prioritised replay, batch 512*2,
            ring of 100 rows x 512 envs (51k transitions = 2.9 GB uint8 x2; the full 1M = 56 GB also fits)
Prints env-steps/s of collection, updates/s, and combined time per epoch (CUDA events around whole epochs).
"""
import json, os, sys, time
import numpy as np, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchrl_b200.networks as networks, torchrl_b200.policies as policies
from torchrl_b200.algo import TwinSACQ, QRDQN
from torchrl_b200.collector import VecCollector, PixelVecCollector
from torchrl_b200.env import get_vec_env
from torchrl_b200.replay_buffers import BaseReplayBuffer, PrioritizedReplayBuffer
from torchrl_b200.utils import NullLogger
dev = torch.device("cuda:0")

def timeit(col, agent, epochs=3):
    for e in range(2):
        agent.current_epoch = e; col.train_one_epoch(); agent.update_per_epoch()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tc = tu = 0.0
    for e in range(epochs):
        ev[0].record(); col.rollout_no_sync(); ev[1].record(); agent.update_per_epoch(flush_infos=False); ev[2].record()
        torch.cuda.synchronize()
        tc += ev[0].elapsed_time(ev[1]); tu += ev[1].elapsed_time(ev[2])
    return tc / epochs, tu / epochs

# ---- config 3
N = 1024
env = get_vec_env("SynthAnt-v0", {"reward_scale": 1, "obs_norm": False}, N); ev_env = get_vec_env("SynthAnt-v0", {"obs_norm": False}, N)
env.seed(0); torch.manual_seed(0); np.random.seed(0)
buf = BaseReplayBuffer(env_nums=N, max_replay_buffer_size=int(1e6))
net = dict(hidden_shapes=[256, 256], append_hidden_shapes=[], base_type=networks.MLPBase, activation_func=nn.ReLU)
pf = policies.GuassianContPolicy(input_shape=111, output_shape=16, tanh_action=True, **net)
qf1 = networks.QNet(input_shape=119, output_shape=1, **net); qf2 = networks.QNet(input_shape=119, output_shape=1, **net)
col = VecCollector(env=env, eval_env=ev_env, pf=pf, replay_buffer=buf, device=dev, epoch_frames=64 * N, max_episode_frames=999)
agent = TwinSACQ(pf=pf, qf1=qf1, qf2=qf2, plr=3e-4, qlr=3e-4, policy_std_reg_weight=0, policy_mean_reg_weight=0,
                 env=env, replay_buffer=buf, collector=col, logger=NullLogger(), discount=0.99, batch_size=4 * N, device=dev,
                 save_dir=None, tau=0.005, opt_times=64, num_epochs=10)
tc, tu = timeit(col, agent)
print(json.dumps({"config": 3, "algo": "TwinSACQ", "envs": N, "collect_env_steps_per_s": 64 * N / tc * 1e3,
                  "updates_per_s": 64 / tu * 1e3, "batch": 4 * N, "ms_collect_64_steps": tc, "ms_64_updates": tu,
                  "env_steps_per_s_at_1_update_per_step": 64 * N / (tc + tu) * 1e3}), flush=True)
del agent, col, buf, env, ev_env; torch.cuda.empty_cache()

# ---- config 4
N = 512
env = get_vec_env("SynthAtari-v0", {}, N); ev_env = get_vec_env("SynthAtari-v0", {}, N)
env.seed(0); torch.manual_seed(0); np.random.seed(0)
buf = PrioritizedReplayBuffer(env_nums=N, max_replay_buffer_size=100 * N)
Q = 200
qf = networks.Net(input_shape=(4, 84, 84), output_shape=6 * Q,
                  hidden_shapes=[[16, [8, 8], [4, 4], [0, 0]], [32, [4, 4], [2, 2], [0, 0]], [64, [3, 3], [1, 1], [0, 0]]],
                  append_hidden_shapes=[512], base_type=networks.CNNBase, activation_func=nn.ReLU)
pf = policies.EpsilonGreedyQRDQNDiscretePolicy(quantile_num=Q, qf=qf, start_epsilon=0.1, end_epsilon=0.1, decay_frames=1000000, action_shape=6)
col = PixelVecCollector(env=env, eval_env=ev_env, pf=pf, replay_buffer=buf, device=dev, epoch_frames=32 * N, max_episode_frames=50000)
agent = QRDQN(quantile_num=Q, qf=qf, pf=pf, qlr=5e-5, optimizer_info={"eps": 0.0003125}, env=env, replay_buffer=buf, collector=col,
              logger=NullLogger(), discount=0.99, batch_size=2 * N, device=dev, save_dir=None, opt_times=16,
              use_soft_update=False, target_hard_update_period=10000, num_epochs=10)
tc, tu = timeit(col, agent)
# prioritised sampling + priority update microbench at the config's ring size
u = torch.rand(2, dtype=torch.float64, device=dev)
from torchrl_b200 import ops
buf._priorities[:100] = torch.rand(100, device=dev) + 0.01
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(100): ops.per_sample(buf._priorities, 100, u, 0.4)
e.record(); torch.cuda.synchronize()
print(json.dumps({"config": 4, "algo": "QRDQN(200 quantiles)+PER", "envs": N, "collect_env_steps_per_s": 32 * N / tc * 1e3,
                  "updates_per_s": 16 / tu * 1e3, "batch": 2 * N, "ms_collect_32_steps": tc, "ms_16_updates": tu,
                  "per_sample_us": s.elapsed_time(e) * 10}), flush=True)
