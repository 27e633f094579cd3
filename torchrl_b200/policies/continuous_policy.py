"""Continuous-action policies (API of /root/reference/torchrl/policies/continuous_policy.py).

Same class names (including the reference's "Guassian" spelling), constructor kwargs and
returned dict keys.  The network forward stays in PyTorch; sampling / log-probabilities /
entropy are produced by the CUDA library in one launch (csrc/collect.cu) and nothing is
copied to the host inside ``explore``.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn

from .. import networks
from .. import ops
from . import distribution as D

LOG_SIG_MAX = 2
LOG_SIG_MIN = -20
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class _DeviceRng:
    """Philox stream state shared by a policy's sampling launches: (seed, device counter)."""

    def __init__(self):
        self.seed = None
        self.counter = None

    def ensure(self, device):
        if self.counter is None or self.counter.device != torch.device(device):
            # one stream per rank: shards of a data-parallel job must not explore with identical noise
            rank = int(os.environ.get("RANK", "0"))
            self.seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * rank) & ((1 << 63) - 1)
            self.counter = torch.zeros(1, dtype=torch.int64, device=device)
        return self


class UniformPolicyContinuous(nn.Module):
    def __init__(self, action_shape):
        super().__init__()
        self.continuous = True
        self.action_shape = action_shape

    def forward(self, x):
        return torch.Tensor(np.random.uniform(-1., 1., self.action_shape))

    def explore(self, x):
        return {"action": torch.Tensor(np.random.uniform(-1., 1., self.action_shape)).to(x.device)}


class DetContPolicy(networks.Net):
    def __init__(self, tanh_action=False, **kwargs):
        super().__init__(**kwargs)
        self.continuous = True
        self.tanh_action = tanh_action

    def forward(self, x):
        out = super().forward(x)
        return torch.tanh(out) if self.tanh_action else out

    def eval_act(self, x):
        with torch.no_grad():
            return self.forward(x).squeeze(0)

    def explore(self, x):
        return {"action": self.forward(x).squeeze(0)}


class FixGuassianContPolicy(networks.Net):
    """Deterministic net + fixed-std Gaussian exploration noise (TD3/DDPG; continuous_policy.py:50-74)."""

    def __init__(self, norm_std_explore, tanh_action=False, **kwargs):
        super().__init__(**kwargs)
        self.continuous = True
        self.tanh_action = tanh_action
        self.norm_std_explore = norm_std_explore
        self._rng = _DeviceRng()

    def forward(self, x):
        out = super().forward(x)
        return torch.tanh(out) if self.tanh_action else out

    def eval_act(self, x):
        with torch.no_grad():
            return self.forward(x).squeeze(0)

    def explore(self, x):
        action = self.forward(x).squeeze(0)
        if D.get_noise_mode() == "reference_cpu":
            # Normal(0, std).sample(shape) on the CPU generator, like the reference
            noise = torch.distributions.Normal(0, self.norm_std_explore).sample(action.shape).to(action.device)
            return {"action": action + noise}
        rng = self._rng.ensure(action.device)
        zero_ls = torch.zeros(action.shape[-1], device=action.device)
        out = ops.tanh_gaussian_sample(action.detach().contiguous(), zero_ls, tanh_action=False,
                                       noise_scale=float(self.norm_std_explore), rng=rng)
        ops.counter_advance(rng.counter)
        noise = out["action"] - action.detach()            # sigma * N(0,1), Philox
        return {"action": action + noise}


class GuassianContPolicyBase:
    """explore / update / eval_act shared by the Gaussian policies (continuous_policy.py:77-153)."""

    def _rng_state(self, device):
        if not hasattr(self, "_rng"):
            self._rng = _DeviceRng()
        return self._rng.ensure(device)

    def eval_act(self, x):
        with torch.no_grad():
            mean, _, _ = self.forward(x)
        if self.tanh_action:
            mean = torch.tanh(mean)
        return mean.squeeze(0).detach()

    def torch_eval_act(self, x):
        with torch.no_grad():
            mean, _, _ = self.forward(x)
        if self.tanh_action:
            mean = torch.tanh(mean)
        return mean.detach()

    def explore(self, x, return_log_probs=False, return_pre_tanh=False, eps=None):
        """Sample an action.  Returns the reference's dict: mean, log_std, std, ent, action
        [, log_prob, pre_tanh].  `eps` optionally supplies the N(0,1) noise (tests)."""
        mean, std, log_std = self.forward(x)
        mean = mean if mean.is_contiguous() else mean.contiguous()
        ls = log_std if log_std.dim() == 1 else log_std.expand_as(mean).contiguous()
        if eps is None and D.get_noise_mode() == "reference_cpu":
            eps = D.draw_reference_noise(tuple(mean.shape), mean.device)
        rng = self._rng_state(mean.device)
        need_grad = torch.is_grad_enabled() and (mean.requires_grad or ls.requires_grad)
        if need_grad:
            action, log_prob, z = D._SampleFn.apply(mean, ls, eps, bool(self.tanh_action), bool(return_log_probs), rng)
        else:
            out = ops.tanh_gaussian_sample(mean.detach(), ls.detach(), eps=eps, tanh_action=bool(self.tanh_action),
                                           want_log_prob=return_log_probs, want_pre_tanh=return_pre_tanh, rng=rng)
            action, log_prob, z = out["action"], out.get("log_prob"), out.get("pre_tanh")
        if eps is None:
            ops.counter_advance(rng.counter)
        ent = (0.5 + _HALF_LOG_2PI + log_std).expand_as(mean).sum(-1, keepdim=True)
        dic = {"mean": mean, "log_std": log_std, "std": std, "ent": ent}
        if return_log_probs:
            dic["log_prob"] = log_prob
        if (return_log_probs or return_pre_tanh) and self.tanh_action and z is not None:
            dic["pre_tanh"] = z.squeeze(0)
        dic["action"] = action.squeeze(0)
        return dic

    def act_only(self, x, eps=None, action_out=None, nan_flag=None):
        """Collector fast path: sampled action only (no entropy / dict), one launch after the MLP."""
        if hasattr(self, "mean_and_log_std"):
            mean, log_std = self.mean_and_log_std(x)         # no exp(log_std) launch: the sampler works on log_std
        else:
            mean, _, log_std = self.forward(x)
        mean = mean if mean.is_contiguous() else mean.contiguous()
        ls = log_std if log_std.dim() == 1 else log_std.expand_as(mean).contiguous()
        rng = self._rng_state(mean.device)
        out = ops.tanh_gaussian_sample(mean, ls, eps=eps, tanh_action=bool(self.tanh_action), rng=rng,
                                       nan_flag=nan_flag, action_out=action_out)
        if eps is None:
            ops.counter_advance(rng.counter)
        return out["action"]

    def update(self, obs, actions):
        """log-prob / entropy of given actions with autograd (cold path; the PPO hot path uses
        ops.ppo_actor_loss directly)."""
        mean, std, log_std = self.forward(obs)
        if self.tanh_action:
            dis = D.TanhNormal(mean, std)
            log_prob = dis.log_prob(actions).sum(-1, keepdim=True)
        else:
            log_prob = (-((actions - mean) ** 2) / (2 * std ** 2) - torch.log(std) - _HALF_LOG_2PI).sum(-1, keepdim=True)
        ent = (0.5 + _HALF_LOG_2PI + torch.log(std)).sum(-1, keepdim=True)
        return {"mean": mean, "dis": torch.distributions.Normal(mean, std, validate_args=False), "log_std": log_std, "std": std,
                "log_prob": log_prob, "ent": ent}


class GuassianContPolicy(networks.Net, GuassianContPolicyBase):
    """State-dependent std: the net outputs [mean | log_std] (continuous_policy.py:156-170)."""

    def __init__(self, tanh_action=False, **kwargs):
        super().__init__(**kwargs)
        self.continuous = True
        self.tanh_action = tanh_action

    def forward(self, x):
        out = super().forward(x)
        mean, log_std = out.chunk(2, dim=-1)
        log_std = torch.clamp(log_std, LOG_SIG_MIN, LOG_SIG_MAX)
        return mean, torch.exp(log_std), log_std


class GuassianContPolicyBasicBias(networks.Net, GuassianContPolicyBase):
    """Mean net + a free log-std parameter initialised to log(log_init) (continuous_policy.py:173-188)."""

    def __init__(self, output_shape, tanh_action=False, log_init=0.125, **kwargs):
        super().__init__(output_shape=output_shape, **kwargs)
        self.continuous = True
        self.logstd = nn.Parameter(torch.ones(output_shape) * np.log(log_init))
        self.tanh_action = tanh_action

    def mean_net(self, x):
        return networks.Net.forward(self, x)

    def mean_params(self):
        """Parameters of the mean network only (everything but the free log-std)."""
        return [p for n, p in self.named_parameters() if n != "logstd"]

    def clamped_logstd(self):
        return torch.clamp(self.logstd, LOG_SIG_MIN, LOG_SIG_MAX)

    def mean_and_log_std(self, x):
        """forward() without the std = exp(log_std) tensor (callers that only sample)."""
        return networks.Net.forward(self, x), torch.clamp(self.logstd, LOG_SIG_MIN, LOG_SIG_MAX)

    def forward(self, x):
        mean = networks.Net.forward(self, x)
        logstd = torch.clamp(self.logstd, LOG_SIG_MIN, LOG_SIG_MAX)
        std = torch.exp(logstd).unsqueeze(0).expand_as(mean)
        return mean, std, logstd
