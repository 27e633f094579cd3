"""Advantage actor-critic (API of /root/reference/torchrl/algo/on_policy/a2c.py:8-112).

The constructor is shared with PPO: it re-homes pf and vf into one flat buffer with a fused
clip(0.5)+Adam(eps=1e-5) step per network (the reference builds two torch.optim.Adam with
eps=1e-5 and clips each net's gradient to 0.5, a2c.py:29-39,86-93).
"""
import torch
import torch.optim as optim

from ...flat import FlatAdam
from ..rl_algo import SegmentOptimizer
from .on_rl_algo import OnRLAlgo


class A2C(OnRLAlgo):
    def __init__(self, pf, vf, plr=3e-4, vlr=3e-4, optimizer_class=optim.Adam, entropy_coeff=0.001, **kwargs):
        super().__init__(**kwargs)
        self.pf = pf
        self.vf = vf
        self.to(self.device)
        self.plr = plr
        self.vlr = vlr
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("torchrl_b200 fuses clip+Adam in CUDA; only optim.Adam is supported")
        self.optimizer_class = optimizer_class
        # segment 0 = policy, segment 1 = value net
        self.opt = FlatAdam([self.pf, self.vf], lrs=[plr, vlr], eps=1e-5, max_norms=[0.5, 0.5], device=self.device)
        self.pf_optimizer = SegmentOptimizer(self.opt, 0)
        self.vf_optimizer = SegmentOptimizer(self.opt, 1)
        self.entropy_coeff = entropy_coeff
        self.vf_criterion = torch.nn.MSELoss()

    def update(self, batch):
        """One A2C minibatch update (a2c.py:45-112); torch autograd for the losses (cold path),
        fused clip+Adam for the step."""
        self.training_update_num += 1
        dev = self.device
        obs = torch.as_tensor(batch['obs'], dtype=torch.float32, device=dev)
        acts = torch.as_tensor(batch['acts'], dtype=torch.float32, device=dev)
        advs = torch.as_tensor(batch['advs'], dtype=torch.float32, device=dev)
        est_rets = torch.as_tensor(batch['estimate_returns'], dtype=torch.float32, device=dev)
        out = self.pf.update(obs, acts)
        log_probs, ent = out['log_prob'], out['ent']
        advs = (advs - advs.mean()) / (advs.std() + 1e-5)
        assert log_probs.shape == advs.shape
        policy_loss = (-log_probs * advs).mean() - self.entropy_coeff * ent.mean()
        values = self.vf(obs)
        vf_loss = self.vf_criterion(values, est_rets)
        (policy_loss + vf_loss).backward()      # disjoint parameter sets: same grads as two backward calls
        self.opt.step()
        info = {'Training/policy_loss': policy_loss.item(), 'Training/vf_loss': vf_loss.item(),
                'v_pred/mean': values.mean().item(), 'v_pred/std': values.std().item(),
                'v_pred/max': values.max().item(), 'v_pred/min': values.min().item()}
        if 'std' in out:
            std = out['std']
            info.update({'std/mean': std.mean().item(), 'std/std': std.std().item(),
                         'std/max': std.max().item(), 'std/min': std.min().item()})
        info['ent'] = ent.mean().item()
        info['log_prob'] = log_probs.mean().item()
        return info

    @property
    def snapshot_networks(self):
        return [("pf", self.pf), ("vf", self.vf)]
