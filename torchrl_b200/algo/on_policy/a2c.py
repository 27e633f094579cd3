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
        self.opt = FlatAdam([self.pf, self.vf], lrs=[plr, vlr], eps=1e-5, max_norms=[0.5, 0.5], device=self.device, dist=self.dist)
        self.pf_optimizer = SegmentOptimizer(self.opt, 0)
        self.vf_optimizer = SegmentOptimizer(self.opt, 1)
        self.entropy_coeff = entropy_coeff
        self.vf_criterion = torch.nn.MSELoss()

    def _minibatch(self, batch, keys):
        return [torch.as_tensor(batch[k], dtype=torch.float32, device=self.device) for k in keys]

    @staticmethod
    def _four_stats(prefix, t):
        return {prefix + '/mean': t.mean().item(), prefix + '/std': t.std().item(),
                prefix + '/max': t.max().item(), prefix + '/min': t.min().item()}

    def update(self, batch):
        """One A2C minibatch update (a2c.py:45-112): policy-gradient loss with normalised advantages and an
        entropy bonus, MSE critic; torch autograd for the two losses (cold path; PPO is the tuned agent), fused
        clip(0.5)+Adam for both networks in one step."""
        self.training_update_num += 1
        obs, acts, advs, est_rets = self._minibatch(batch, ('obs', 'acts', 'advs', 'estimate_returns'))
        dist_out = self.pf.update(obs, acts)
        log_probs, ent = dist_out['log_prob'], dist_out['ent']
        advs = (advs - advs.mean()) / (advs.std() + 1e-5)
        assert log_probs.shape == advs.shape, (log_probs.shape, advs.shape)
        policy_loss = -(log_probs * advs).mean() - self.entropy_coeff * ent.mean()
        values = self.vf(obs)
        vf_loss = self.vf_criterion(values, est_rets)
        (policy_loss + vf_loss).backward()      # disjoint parameter sets: same grads as two backward calls
        self.opt.step()
        info = {'Training/policy_loss': policy_loss.item(), 'Training/vf_loss': vf_loss.item()}
        info.update(self._four_stats('v_pred', values))
        if 'std' in dist_out:
            info.update(self._four_stats('std', dist_out['std']))
        info['ent'] = ent.mean().item()
        info['log_prob'] = log_probs.mean().item()
        return info

    @property
    def snapshot_networks(self):
        return [("pf", self.pf), ("vf", self.vf)]
