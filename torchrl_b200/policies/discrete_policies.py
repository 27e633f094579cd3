"""Discrete-action policies (API of /root/reference/torchrl/policies/discrete_policies.py).

Differences from the reference, all documented in SURVEY.md Appendix A: the QR-DQN policy's
`q_to_a` works for any number of envs (the reference calls .item(), A.5) and epsilon-greedy
draws stay on the device unless the noise mode is "reference_cpu" (then np.random is used in
the reference's call order).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.distributions import Categorical

from .. import networks
from . import distribution as D


class UniformPolicyDiscrete(nn.Module):
    def __init__(self, action_num):
        super().__init__()
        self.action_num = action_num
        self.continuous = False

    def forward(self, x):
        return np.random.randint(self.action_num)

    def explore(self, x):
        return {"action": np.random.randint(self.action_num)}


def linear_decay(start, end, frames, count):
    """Exploration rate after `count` decisions: linear from `start` to `end` over `frames`, then flat
    (discrete_policies.py:44-49)."""
    if count >= frames:
        return end
    return start - (start - end) * (count / frames)


def _uniform_and_random_actions(shape, n_actions, device):
    """(u in [0,1), a in {0..n_actions-1}) per decision.  "reference_cpu" mode draws them from the global NumPy
    stream in the reference's order (rand, then randint); otherwise they are generated on the device."""
    if D.get_noise_mode() == "reference_cpu":
        u = torch.Tensor(np.random.rand(*shape)).to(device)
        a = torch.LongTensor(np.random.randint(low=0, high=n_actions, size=shape)).to(device)
        return u, a
    return torch.rand(shape, device=device), torch.randint(0, n_actions, shape, device=device)


class EpsilonGreedyDQNDiscretePolicy:
    """epsilon-greedy decisions on top of a Q network (discrete_policies.py:25-74).  Not an nn.Module, like in the
    reference: `to` / `parameters` forward to the wrapped network."""

    def __init__(self, qf, start_epsilon, end_epsilon, decay_frames, action_shape):
        self.qf = qf
        self.start_epsilon, self.end_epsilon, self.decay_frames = start_epsilon, end_epsilon, decay_frames
        self.action_shape = action_shape
        self.count = 0
        self.epsilon = start_epsilon
        self.continuous = False

    def q_to_a(self, q):
        return q.max(dim=-1, keepdim=True)[1].detach()

    def tick(self):
        """Advance the schedule by one decision (host side).  Collectors that replay a captured step graph call
        this outside the graph and hand the rate over in a device scalar (`explore(x, epsilon=tensor)`)."""
        self.count += 1
        self.epsilon = linear_decay(self.start_epsilon, self.end_epsilon, self.decay_frames, self.count)

    def explore(self, x, epsilon=None):
        if epsilon is None:
            self.tick()
            epsilon = self.epsilon
        x = x.squeeze(0)
        output = self.qf(x)
        greedy = self.q_to_a(output)
        u, random_action = _uniform_and_random_actions(tuple(greedy.shape), self.action_shape, x.device)
        return {"q_value": output, "action": torch.where(u < epsilon, random_action, greedy)}

    def eval_act(self, x):
        with torch.no_grad():
            return self.q_to_a(self.qf(x))

    def to(self, device):
        self.qf.to(device)
        return self

    def parameters(self):
        return self.qf.parameters()


class EpsilonGreedyQRDQNDiscretePolicy(EpsilonGreedyDQNDiscretePolicy):
    """Greedy w.r.t. the mean over quantiles (discrete_policies.py:77-89), batched."""

    def __init__(self, quantile_num, **kwargs):
        super().__init__(**kwargs)
        self.quantile_num = quantile_num
        self.continuous = False

    def q_to_a(self, q):
        q = q.view(q.shape[:-1] + (self.action_shape, self.quantile_num))
        return q.mean(dim=-1).max(dim=-1, keepdim=True)[1].detach()


class CategoricalDisPolicy(networks.Net):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.continuous = False

    def forward(self, x):
        return torch.softmax(super().forward(x), dim=-1)

    def explore(self, x, return_log_probs=False):
        output = self.forward(x)
        dis = Categorical(output)
        action = dis.sample()
        out = {"dis": output, "action": action}
        if return_log_probs:
            out["log_prob"] = dis.log_prob(action)
        return out

    def eval_act(self, x):
        return self.forward(x).max(dim=-1)[1].detach()

    def update(self, obs, actions):
        dis = Categorical(self.forward(obs))
        return {"dis": dis, "log_prob": dis.log_prob(actions).unsqueeze(-1), "ent": dis.entropy()}
