"""CPU port of the reference's PPO / TwinSAC-Q / TD3 hot path (TEST + BASELINE INFRASTRUCTURE).

A from-scratch restatement, in one file, of what the reference does on the CPU for the
north-star path -- per-env Python wrappers, in-process and multi-process vectorised envs
(one Pipe per worker, pickled messages, spawn), float64 time-major NumPy buffers, the Python
GAE loop, row-granular minibatches, op-by-op torch updates with a host sync per logged scalar.
It exists because the reference itself (pure Python under /root/reference) cannot travel to
the GPU box: there, ``bench.py --impl reference`` and the ``cpu_baseline`` leg time THIS port.
In the build container tests/test_oracle_vs_reference.py runs the unmodified reference and this
port on identical seeds and requires identical results, which is what pins the port.

Nothing in the product (torchrl_b200/) imports this file.

Reference lines restated (paths under /root/reference/torchrl):
  env/continuous_wrapper.py:7-20 (NormAct) · env/base_wrapper.py:32-41 (RewardShift), :44-100
  (Normalizer), :103-121 (NormObs), :151-159 (TimeLimitAugment) · env/get_env.py:52-87 ·
  env/vecenv.py:6-78 · env/subproc_vecenv.py:10-157 · replay_buffers/base.py:4-54 ·
  replay_buffers/on_policy.py:5-95 · networks/init.py, base.py:8-44, nets.py:13-68 ·
  policies/continuous_policy.py:77-188 · policies/distribution.py:5-79 ·
  collector/base.py:10-230, collector/on_policy.py:84-155 · algo/on_policy/{on_rl_algo,a2c,ppo}.py ·
  algo/off_policy/{off_rl_algo,twin_sac_q,td3}.py · algo/utils.py
"""
import copy
import math
import os
import sys

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import synth_env  # noqa: E402
from oracle.ref_numpy import RunningNorm  # noqa: E402


# ======================================================================================= envs
from oracle.cpu_envs import (WrappedEnv, InProcVecEnv, SubprocVecEnv, NormObsVec, make_vec_env)  # noqa: E402,F401


# ======================================================================================= buffers
class Float64Rollout:
    """Time-major float64 ring (replay_buffers/base.py + on_policy.py)."""

    def __init__(self, max_size, env_nums=1, time_limit_filter=False):
        self.env_nums = env_nums
        self.rows = max_size // env_nums
        self.top, self.size = 0, 0
        self.time_limit_filter = time_limit_filter
        self.data = {}

    def add(self, sample):
        for k, v in sample.items():
            if k not in self.data:
                self.data[k] = np.zeros((self.rows,) + np.shape(v))
            self.data[k][self.top, ...] = v
        self.top = (self.top + 1) % self.rows
        self.size = min(self.size + 1, self.rows)

    def last(self, keys):
        return {k: self.data[k][self.rows - 1] for k in keys}

    def gae(self, last_value, gamma, tau):
        """The reference's Python loop with list.insert(0, ...) (on_policy.py:16-44)."""
        d = self.data
        run = 0
        advs, rets = [], []
        values = np.concatenate([d["values"], np.array([last_value])], 0)
        for t in reversed(range(len(d["rewards"]))):
            delta = d["rewards"][t] + (1 - d["terminals"][t]) * gamma * values[t + 1] - values[t]
            run = delta + (1 - d["terminals"][t]) * gamma * tau * run
            if self.time_limit_filter:
                run = run * (1 - d["time_limits"][t])
            advs.insert(0, run)
            rets.insert(0, run + values[t])
        d["advs"] = np.array(advs)
        d["estimate_returns"] = np.array(rets)

    def minibatches(self, batch_size, keys, shuffle):
        assert batch_size % self.env_nums == 0
        b = batch_size // self.env_nums
        order = np.random.permutation(self.rows) if shuffle else np.arange(self.rows)
        pos = 0
        while pos < self.rows:
            out = {}
            for k in keys:
                chunk = self.data[k][order[pos:pos + b]]
                out[k] = chunk.reshape((b * self.env_nums,) + chunk.shape[2:])
            yield out
            pos += b

    def random_batch(self, batch_size, keys):
        assert batch_size % self.env_nums == 0
        b = batch_size // self.env_nums
        idx = np.random.randint(0, self.size, b)
        out = {}
        for k in keys:
            chunk = self.data[k][idx]
            out[k] = chunk.reshape((b * self.env_nums,) + chunk.shape[2:])
        return out


# ======================================================================================= nets
def _init_hidden(fc):
    bound = math.sqrt(1.0 / fc.weight.size(0))           # the reference's "fan-in" = size(0) (init.py:5-15)
    fc.weight.data.uniform_(-bound, bound)
    fc.bias.data.fill_(0.1)


def _init_last(fc):
    fc.weight.data.uniform_(-3e-3, 3e-3)
    fc.bias.data.uniform_(-3e-3, 3e-3)


class MLPNet(nn.Module):
    """Net(base_type=MLPBase) (networks/base.py:8-44, nets.py:13-52): hidden Linear+act ... + linear head.
    Parameter creation order and RNG consumption match the reference so equal seeds give equal weights."""

    def __init__(self, in_dim, out_dim, hidden, act=nn.Tanh):
        super().__init__()
        layers, w = [], int(in_dim)
        for h in hidden:
            fc = nn.Linear(w, h)
            _init_hidden(fc)
            layers += [fc, act()]
            w = h
        self.base = nn.Sequential(*layers)
        self.head = nn.Linear(w, out_dim)
        _init_last(self.head)

    def forward(self, x):
        return self.head(self.base(x))


class QNet(MLPNet):
    def forward(self, pair):
        return super().forward(torch.cat(pair, dim=-1))


_HALF_LOG_2PI = 0.5 * math.log(2 * math.pi)


def _cpu_noise(shape):
    """The reference draws exploration noise on the CPU generator (distribution.py:64-70)."""
    return torch.distributions.Normal(torch.zeros(shape), torch.ones(shape)).sample()


class TanhGaussianPolicy(nn.Module):
    """GuassianContPolicyBasicBias (shared free log-std) or GuassianContPolicy (state-dependent)."""

    def __init__(self, in_dim, act_dim, hidden, act=nn.Tanh, state_dependent_std=False, log_init=0.125):
        super().__init__()
        self.act_dim = act_dim
        self.state_dependent_std = state_dependent_std
        self.net = MLPNet(in_dim, act_dim * (2 if state_dependent_std else 1), hidden, act)
        if not state_dependent_std:
            self.logstd = nn.Parameter(torch.ones(act_dim) * np.log(log_init))

    def dist_params(self, x):
        out = self.net(x)
        if self.state_dependent_std:
            mean, log_std = out.chunk(2, dim=-1)
            log_std = torch.clamp(log_std, -20, 2)
            return mean, torch.exp(log_std), log_std
        log_std = torch.clamp(self.logstd, -20, 2)
        return out, torch.exp(log_std).unsqueeze(0).expand_as(out), log_std

    def explore(self, x, return_log_probs=False):
        mean, std, log_std = self.dist_params(x)
        ent = (0.5 + _HALF_LOG_2PI + torch.log(std)).sum(-1, keepdim=True)
        z = mean + std * _cpu_noise(mean.size()).to(mean.device)
        action = torch.tanh(z)
        out = {"mean": mean, "log_std": log_std, "std": std, "ent": ent}
        if return_log_probs:
            lp = torch.distributions.Normal(mean, std).log_prob(z) - torch.log(1 - action * action + 1e-6)
            out["log_prob"] = lp.sum(dim=-1, keepdim=True)
        out["action"] = action.squeeze(0)
        return out

    def evaluate(self, obs, actions):
        """log-prob of stored actions through atanh recomputation, and Normal entropy."""
        mean, std, log_std = self.dist_params(obs)
        pre = torch.log((1 + actions) / (1 - actions)) / 2
        lp = torch.distributions.Normal(mean, std).log_prob(pre) - torch.log(1 - actions * actions + 1e-6)
        ent = torch.distributions.Normal(mean, std).entropy().sum(-1, keepdim=True)
        return {"log_prob": lp.sum(-1, keepdim=True), "ent": ent, "log_std": log_std, "mean": mean, "std": std}

    def eval_act(self, x):
        with torch.no_grad():
            mean, _, _ = self.dist_params(x)
        return torch.tanh(mean).squeeze(0).numpy()


class FixedNoisePolicy(nn.Module):
    """FixGuassianContPolicy (TD3): deterministic net + N(0, std) exploration noise."""

    def __init__(self, in_dim, act_dim, hidden, act=nn.ReLU, norm_std_explore=0.1, tanh_action=True):
        super().__init__()
        self.net = MLPNet(in_dim, act_dim, hidden, act)
        self.norm_std_explore = norm_std_explore
        self.tanh_action = tanh_action

    def forward(self, x):
        out = self.net(x)
        return torch.tanh(out) if self.tanh_action else out

    def explore(self, x):
        action = self.forward(x).squeeze(0)
        noise = torch.distributions.Normal(0, self.norm_std_explore).sample(action.shape)
        return {"action": action + noise}


# ======================================================================================= collectors
class OnPolicyVecCollector:
    """VecOnPolicyCollector.take_actions / train_one_epoch (collector/on_policy.py:84-155, base.py:108-122)."""

    def __init__(self, env, pf, vf, buffer, epoch_frames, max_episode_frames=999, discount=0.99):
        self.env, self.pf, self.vf, self.buffer = env, pf, vf, buffer
        self.env.train()
        self.current_ob = self.env.reset()
        self.steps_per_epoch = epoch_frames // env.env_nums
        self.max_episode_frames = max_episode_frames
        self.discount = discount
        self.current_step = np.zeros((env.env_nums, 1))
        self.train_rew = np.zeros_like(self.current_step)

    def step(self):
        ob_t = torch.Tensor(self.current_ob)
        acts = self.pf.explore(ob_t)["action"].detach().cpu().numpy()
        values = self.vf(ob_t).detach().cpu().numpy()
        if np.isnan(acts).any():
            raise FloatingPointError("NaN detected. BOOM")
        next_obs, rewards, dones, infos = self.env.step(acts)
        self.current_step += 1
        sample = {"obs": self.current_ob, "next_obs": next_obs, "acts": acts, "values": values, "rewards": rewards,
                  "terminals": dones, "time_limits": infos["time_limit"][:, np.newaxis] if "time_limit" in infos
                  else [False]}
        self.train_rew += rewards
        if np.any(dones):
            self.train_rews += list(self.train_rew[dones])
            self.train_rew[dones] = 0
        if np.any(dones) or np.any(self.current_step >= self.max_episode_frames):
            surpass = self.current_step >= self.max_episode_frames
            last_value = self.vf(torch.Tensor(next_obs)).detach().cpu().numpy()
            sample["terminals"] = dones | surpass
            sample["rewards"] = rewards + self.discount * last_value * surpass
            next_obs = self.env.partial_reset(np.squeeze(dones | surpass, axis=-1))
            self.current_step[dones | surpass] = 0
        self.buffer.add(sample)
        self.current_ob = next_obs
        return np.sum(rewards)

    def train_one_epoch(self):
        self.train_rews = []
        total = 0
        self.env.train()
        for _ in range(self.steps_per_epoch):
            total += self.step()
        return {"train_rewards": self.train_rews, "train_epoch_reward": total}


class OffPolicyVecCollector:
    """VecCollector.take_actions (collector/base.py:184-230): no values, reset on done or timeout."""

    def __init__(self, env, pf, buffer, epoch_frames, max_episode_frames=999):
        self.env, self.pf, self.buffer = env, pf, buffer
        self.env.train()
        self.current_ob = self.env.reset()
        self.steps_per_epoch = epoch_frames // env.env_nums
        self.max_episode_frames = max_episode_frames
        self.current_step = np.zeros((env.env_nums, 1))
        self.train_rew = np.zeros_like(self.current_step)

    def step(self):
        out = self.pf.explore(torch.Tensor(self.current_ob).unsqueeze(0))
        act = out["action"].detach().cpu().numpy()
        if np.isnan(act).any():
            raise FloatingPointError("NaN detected. BOOM")
        next_ob, reward, done, infos = self.env.step(act)
        self.current_step += 1
        sample = {"obs": self.current_ob, "next_obs": next_ob, "acts": act, "rewards": reward, "terminals": done,
                  "time_limits": infos["time_limit"][:, np.newaxis] if "time_limit" in infos else [False]}
        self.train_rew += reward
        if np.any(done):
            self.train_rews += list(self.train_rew[done])
            self.train_rew[done] = 0
        if np.any(done) or np.any(self.current_step >= self.max_episode_frames):
            flag = (self.current_step >= self.max_episode_frames) | done
            next_ob = self.env.partial_reset(np.squeeze(flag, axis=-1))
            self.current_step[flag] = 0
        self.buffer.add(sample)
        self.current_ob = next_ob
        return np.sum(reward)

    def train_one_epoch(self):
        self.train_rews = []
        total = 0
        self.env.train()
        for _ in range(self.steps_per_epoch):
            total += self.step()
        return {"train_rewards": self.train_rews, "train_epoch_reward": total}


# ======================================================================================= PPO
class PPOPort:
    """PPO.update_per_epoch / update (algo/on_policy/ppo.py:27-152; a2c.py:29-39 for the optimizers)."""

    def __init__(self, pf, vf, buffer, plr=3e-4, vlr=3e-4, clip_para=0.2, opt_epochs=10, entropy_coeff=0.005,
                 tau=0.95, discount=0.99, batch_size=64, num_epochs=488, shuffle=True, clipped_value_loss=False):
        self.pf, self.vf, self.buffer = pf, vf, buffer
        self.target_pf = copy.deepcopy(pf)
        self.plr, self.vlr = plr, vlr
        self.pf_opt = torch.optim.Adam(pf.parameters(), lr=plr, eps=1e-5)
        self.vf_opt = torch.optim.Adam(vf.parameters(), lr=vlr, eps=1e-5)
        self.clip_para, self.opt_epochs, self.entropy_coeff = clip_para, opt_epochs, entropy_coeff
        self.tau, self.discount, self.batch_size = tau, discount, batch_size
        self.num_epochs, self.shuffle, self.clipped_value_loss = num_epochs, shuffle, clipped_value_loss
        self.current_epoch = 0
        self.infos = []

    def process_epoch_samples(self):
        s = self.buffer.last(["next_obs", "terminals", "time_limits"])
        last_value = self.vf(torch.Tensor(s["next_obs"])).detach().cpu().numpy()
        last_value = last_value * (1 - s["terminals"])
        self.buffer.gae(last_value, self.discount, self.tau)

    def update_per_epoch(self):
        self.process_epoch_samples()
        frac = self.current_epoch / float(self.num_epochs)
        for opt, lr0 in ((self.pf_opt, self.plr), (self.vf_opt, self.vlr)):
            for g in opt.param_groups:
                g["lr"] = lr0 - lr0 * frac
        for tp, p in zip(self.target_pf.parameters(), self.pf.parameters()):
            tp.data.copy_(p.data)
        keys = ["obs", "acts", "advs", "estimate_returns", "values"]
        for _ in range(self.opt_epochs):
            for batch in self.buffer.minibatches(self.batch_size, keys, self.shuffle):
                self.infos.append(self.update(batch))

    def update(self, batch):
        info = {}
        obs = torch.Tensor(batch["obs"])
        actions = torch.Tensor(batch["acts"])
        advs = torch.Tensor(batch["advs"])
        old_values = torch.Tensor(batch["values"])
        est_rets = torch.Tensor(batch["estimate_returns"])
        info["advs/mean"] = advs.mean().item()
        info["advs/std"] = advs.std().item()
        info["advs/max"] = advs.max().item()
        info["advs/min"] = advs.min().item()
        advs = (advs - advs.mean()) / (advs.std() + 1e-5)
        # critic
        values = self.vf(obs)
        if self.clipped_value_loss:
            clipped = old_values + (values - old_values).clamp(-self.clip_para, self.clip_para)
            vf_loss = 0.5 * torch.max((values - est_rets).pow(2), (clipped - est_rets).pow(2)).mean()
        else:
            vf_loss = nn.functional.mse_loss(values, est_rets)
        self.vf_opt.zero_grad()
        vf_loss.backward()
        vf_norm = torch.nn.utils.clip_grad_norm_(self.vf.parameters(), 0.5)
        self.vf_opt.step()
        info["Training/vf_loss"] = vf_loss.item()
        info["grad_norm/vf"] = vf_norm.item()
        # actor
        out = self.pf.evaluate(obs, actions)
        log_probs, ent, log_std = out["log_prob"], out["ent"], out["log_std"]
        with torch.no_grad():
            old_log_probs = self.target_pf.evaluate(obs, actions)["log_prob"]
        ratio = torch.exp(log_probs - old_log_probs.detach())
        s1 = ratio * advs
        s2 = torch.clamp(ratio, 1.0 - self.clip_para, 1.0 + self.clip_para) * advs
        policy_loss = -torch.mean(torch.min(s2, s1)) - self.entropy_coeff * ent.mean()
        self.pf_opt.zero_grad()
        policy_loss.backward()
        pf_norm = torch.nn.utils.clip_grad_norm_(self.pf.parameters(), 0.5)
        self.pf_opt.step()
        info["Training/policy_loss"] = policy_loss.item()
        for name, t in (("logprob", log_probs), ("log_std", log_std)):
            info[name + "/mean"] = t.mean().item()
            info[name + "/std"] = t.std().item()
            info[name + "/max"] = t.max().item()
            info[name + "/min"] = t.min().item()
        info["ratio/max"] = ratio.max().item()
        info["ratio/min"] = ratio.min().item()
        info["grad_norm/pf"] = pf_norm.item()
        return info


def build_ppo(env_id="SynthHalfCheetah-v0", env_nums=8, proc_nums=0, horizon=128, hidden=(256, 256), batch_rows=4,
              opt_epochs=10, seed=0, max_episode_frames=999, obs_norm=True, num_epochs=488, reward_scale=1):
    """Wire the PPO pipeline like examples/ppo_continuous_vec(_subproc).py does (same seeding order)."""
    env = make_vec_env(env_id, {"reward_scale": reward_scale, "obs_norm": obs_norm}, env_nums, proc_nums)
    env.seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    o, a = env.observation_space.shape[0], env.action_space.shape[0]
    buf = Float64Rollout(horizon * env_nums, env_nums, time_limit_filter=True)
    pf = TanhGaussianPolicy(o, a, list(hidden), nn.Tanh)
    vf = MLPNet(o, 1, list(hidden), nn.Tanh)
    col = OnPolicyVecCollector(env, pf, vf, buf, horizon * env_nums, max_episode_frames, 0.99)
    agent = PPOPort(pf, vf, buf, opt_epochs=opt_epochs, batch_size=batch_rows * env_nums, num_epochs=num_epochs)
    return env, col, agent


# ======================================================================================= off-policy
def _polyak(net, target, tau):
    for tp, p in zip(target.parameters(), net.parameters()):
        tp.data.copy_(tp.data * (1.0 - tau) + p.data * tau)


class SACPort:
    """TwinSACQ.update (algo/off_policy/twin_sac_q.py:84-219): automatic temperature, twin critics,
    reparameterised policy loss, three Adam steps, Polyak targets."""

    def __init__(self, pf, qf1, qf2, act_dim, plr=3e-4, qlr=3e-4, discount=0.99, tau=0.005, std_reg=0.0,
                 mean_reg=0.0, grad_clip=None):
        self.pf, self.qf1, self.qf2 = pf, qf1, qf2
        self.tqf1, self.tqf2 = copy.deepcopy(qf1), copy.deepcopy(qf2)
        self.q1_opt = torch.optim.Adam(qf1.parameters(), lr=qlr)
        self.q2_opt = torch.optim.Adam(qf2.parameters(), lr=qlr)
        self.pf_opt = torch.optim.Adam(pf.parameters(), lr=plr)
        self.target_entropy = -float(act_dim)
        self.log_alpha = torch.zeros(1, requires_grad=True)
        self.alpha_opt = torch.optim.Adam([self.log_alpha], lr=plr)
        self.discount, self.tau, self.std_reg, self.mean_reg, self.grad_clip = discount, tau, std_reg, mean_reg, grad_clip

    def update(self, batch):
        rewards = torch.Tensor(batch["rewards"])
        terminals = torch.Tensor(batch["terminals"])
        obs = torch.Tensor(batch["obs"])
        actions = torch.Tensor(batch["acts"])
        next_obs = torch.Tensor(batch["next_obs"])
        s = self.pf.explore(obs, return_log_probs=True)
        mean, log_std, new_actions, log_probs = s["mean"], s["log_std"], s["action"], s["log_prob"]
        q1_pred = self.qf1([obs, actions])
        q2_pred = self.qf2([obs, actions])
        alpha_loss = -(self.log_alpha * (log_probs + self.target_entropy).detach()).mean()
        self.alpha_opt.zero_grad()
        alpha_loss.backward()
        self.alpha_opt.step()
        alpha = self.log_alpha.exp().detach()
        with torch.no_grad():
            t = self.pf.explore(next_obs, return_log_probs=True)
            tq = torch.min(self.tqf1([next_obs, t["action"]]), self.tqf2([next_obs, t["action"]]))
            target_v = tq - alpha * t["log_prob"]
        q_target = rewards + (1.0 - terminals) * self.discount * target_v
        qf1_loss = nn.functional.mse_loss(q1_pred, q_target.detach())
        qf2_loss = nn.functional.mse_loss(q2_pred, q_target.detach())
        q_new = torch.min(self.qf1([obs, new_actions]), self.qf2([obs, new_actions]))
        policy_loss = (alpha * log_probs - q_new).mean()
        policy_loss = policy_loss + self.std_reg * (log_std ** 2).mean() + self.mean_reg * (mean ** 2).mean()
        for opt, loss, net in ((self.pf_opt, policy_loss, self.pf), (self.q1_opt, qf1_loss, self.qf1),
                               (self.q2_opt, qf2_loss, self.qf2)):
            opt.zero_grad()
            loss.backward()
            if self.grad_clip:
                torch.nn.utils.clip_grad_norm_(net.parameters(), self.grad_clip)
            opt.step()
        _polyak(self.qf1, self.tqf1, self.tau)
        _polyak(self.qf2, self.tqf2, self.tau)
        info = {"Reward_Mean": rewards.mean().item(), "Alpha": alpha.item(), "Alpha_loss": alpha_loss.item(),
                "Training/policy_loss": policy_loss.item(), "Training/qf1_loss": qf1_loss.item(),
                "Training/qf2_loss": qf2_loss.item()}
        for name, t_ in (("log_std", log_std), ("log_probs", log_probs), ("mean", mean)):
            info[name + "/mean"] = t_.mean().item()
            info[name + "/std"] = t_.std().item()
            info[name + "/max"] = t_.max().item()
            info[name + "/min"] = t_.min().item()
        return info


class TD3Port:
    """TD3.update (algo/off_policy/td3.py:57-154), including the inverted delay test (:124)."""

    def __init__(self, pf, qf1, qf2, plr=3e-4, qlr=3e-4, discount=0.99, tau=0.005, policy_update_delay=2,
                 norm_std_policy=0.2, noise_clip=0.5):
        self.pf, self.qf1, self.qf2 = pf, qf1, qf2
        self.tpf, self.tqf1, self.tqf2 = copy.deepcopy(pf), copy.deepcopy(qf1), copy.deepcopy(qf2)
        self.pf_opt = torch.optim.Adam(pf.parameters(), lr=plr)
        self.q1_opt = torch.optim.Adam(qf1.parameters(), lr=qlr)
        self.q2_opt = torch.optim.Adam(qf2.parameters(), lr=qlr)
        self.discount, self.tau, self.delay = discount, tau, policy_update_delay
        self.norm_std_policy, self.noise_clip = norm_std_policy, noise_clip
        self.n = 0

    def update(self, batch):
        self.n += 1
        obs = torch.Tensor(batch["obs"])
        actions = torch.Tensor(batch["acts"])
        next_obs = torch.Tensor(batch["next_obs"])
        rewards = torch.Tensor(batch["rewards"])
        terminals = torch.Tensor(batch["terminals"])
        target_actions = self.tpf.explore(next_obs)["action"]
        noise = torch.distributions.Normal(torch.zeros(target_actions.size()),
                                           self.norm_std_policy * torch.ones(target_actions.size())).sample()
        target_actions = target_actions + torch.clamp(noise, -self.noise_clip, self.noise_clip)
        target_actions = torch.clamp(target_actions, -1, 1)
        target_q = torch.min(self.tqf1([next_obs, target_actions]), self.tqf2([next_obs, target_actions]))
        q_target = rewards + (1.0 - terminals) * self.discount * target_q
        q1_pred, q2_pred = self.qf1([obs, actions]), self.qf2([obs, actions])
        qf1_loss = nn.functional.mse_loss(q1_pred, q_target.detach())
        qf2_loss = nn.functional.mse_loss(q2_pred, q_target.detach())
        for opt, loss in ((self.q1_opt, qf1_loss), (self.q2_opt, qf2_loss)):
            opt.zero_grad()
            loss.backward()
            opt.step()
        info = {"Reward_Mean": rewards.mean().item(), "Training/qf1_loss": qf1_loss.item(),
                "Training/qf2_loss": qf2_loss.item()}
        if self.n % self.delay:
            new_actions = self.pf(obs)
            policy_loss = -self.qf1([obs, new_actions]).mean()
            self.pf_opt.zero_grad()
            policy_loss.backward()
            self.pf_opt.step()
            _polyak(self.pf, self.tpf, self.tau)
            _polyak(self.qf1, self.tqf1, self.tau)
            _polyak(self.qf2, self.tqf2, self.tau)
            info["Training/policy_loss"] = policy_loss.item()
            info["new_actions/mean"] = new_actions.mean().item()
            info["new_actions/std"] = new_actions.std().item()
            info["new_actions/max"] = new_actions.max().item()
            info["new_actions/min"] = new_actions.min().item()
        return info


class DDPGPort:
    """DDPG.update (algo/off_policy/ddpg.py:40-111): deterministic policy gradient through the critic, one
    critic regressed on r + (1-d)*gamma*Q'(s', pi'(s')), actor step before critic step, Polyak targets."""

    def __init__(self, pf, qf, plr=3e-4, qlr=3e-4, discount=0.99, tau=0.005):
        self.pf, self.qf = pf, qf
        self.tpf, self.tqf = copy.deepcopy(pf), copy.deepcopy(qf)
        self.pf_opt = torch.optim.Adam(pf.parameters(), lr=plr)
        self.qf_opt = torch.optim.Adam(qf.parameters(), lr=qlr)
        self.discount, self.tau = discount, tau

    def update(self, batch):
        obs = torch.Tensor(batch["obs"])
        actions = torch.Tensor(batch["acts"])
        next_obs = torch.Tensor(batch["next_obs"])
        rewards = torch.Tensor(batch["rewards"])
        terminals = torch.Tensor(batch["terminals"])
        new_actions = self.pf(obs)
        policy_loss = -self.qf([obs, new_actions]).mean()
        target_q = self.tqf([next_obs, self.tpf(next_obs)])
        q_target = rewards + (1.0 - terminals) * self.discount * target_q
        qf_loss = nn.functional.mse_loss(self.qf([obs, actions]), q_target.detach())
        self.pf_opt.zero_grad()
        policy_loss.backward()
        self.pf_opt.step()
        self.qf_opt.zero_grad()
        qf_loss.backward()
        self.qf_opt.step()
        _polyak(self.pf, self.tpf, self.tau)
        _polyak(self.qf, self.tqf, self.tau)
        return {"Reward_Mean": rewards.mean().item(), "Training/policy_loss": policy_loss.item(),
                "Training/qf_loss": qf_loss.item(), "new_actions/mean": new_actions.mean().item(),
                "new_actions/std": new_actions.std().item(), "new_actions/max": new_actions.max().item(),
                "new_actions/min": new_actions.min().item()}
