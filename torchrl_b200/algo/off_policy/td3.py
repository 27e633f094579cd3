"""TD3 (API of /root/reference/torchrl/algo/off_policy/td3.py:10-191).

Keeps the reference's conventions, including the actor/target update firing when
`training_update_num % policy_update_delay` is NON-zero (td3.py:124) and the target
action coming from `target_pf.explore` (exploration noise included) before the clipped smoothing
noise is added (td3.py:72-84).  Two captured graph variants: critics only / critics + actor.
"""
import copy

import torch
import torch.optim as optim

from ... import ops
from ...flat import FlatAdam, FlatParams
from ...policies import distribution as D
from ...policies.continuous_policy import _DeviceRng
from ..rl_algo import SegmentOptimizer
from .off_rl_algo import OffRLAlgo

_STAT = ("mean", "std", "max", "min")


class TD3(OffRLAlgo):
    def __init__(self, pf, qf1, qf2, plr, qlr, optimizer_class=optim.Adam, policy_update_delay=2,
                 norm_std_policy=0.2, noise_clip=0.5, **kwargs):
        super().__init__(**kwargs)
        self.pf = pf
        self.target_pf = copy.deepcopy(pf)
        self.qf1, self.qf2 = qf1, qf2
        self.target_qf1 = copy.deepcopy(qf1)
        self.target_qf2 = copy.deepcopy(qf2)
        self.to(self.device)
        self.plr, self.qlr = plr, qlr
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("torchrl_b200 fuses clip+Adam in CUDA; only optim.Adam is supported")
        clip = self.grad_clip if self.grad_clip else 0.0
        self.opt = FlatAdam([self.pf, self.qf1, self.qf2], lrs=[plr, qlr, qlr], eps=1e-8, max_norms=[clip] * 3,
                            device=self.device, dist=self.dist)
        self.pf_optimizer = SegmentOptimizer(self.opt, 0)
        self.qf1_optimizer = SegmentOptimizer(self.opt, 1)
        self.qf2_optimizer = SegmentOptimizer(self.opt, 2)
        self._target_flat = FlatParams([self.target_pf, self.target_qf1, self.target_qf2], device=self.device)
        self.policy_update_delay = policy_update_delay
        self.norm_std_policy = norm_std_policy
        self.noise_clip = noise_clip
        self._rng = _DeviceRng()

    def _target_source(self):
        return self.opt.data

    def _variant(self):
        return 1 if (self.training_update_num % self.policy_update_delay) else 0

    # info: 0 Reward_Mean | 4 qf1_loss 5 qf2_loss | 6 policy_loss | 10..13 new_actions stats
    def _update_body(self, variant):
        ub = self._ub
        batch = self._batch()
        info = ub["info"][0]
        sc = ub["scratch"]
        obs, acts, next_obs = batch["obs"], batch["acts"], batch["next_obs"]
        rewards, terminals = batch["rewards"].reshape(-1), batch["terminals"].reshape(-1)
        B = obs.shape[0]
        acts = acts.reshape(B, -1)
        with torch.no_grad():
            t_act = self.target_pf.explore(next_obs)["action"].contiguous()
            if D.get_noise_mode() == "reference_cpu":
                eps = torch.distributions.Normal(torch.zeros(t_act.size()), torch.ones(t_act.size())).sample()
                eps = eps.to(t_act.device)
                t_act = ops.td3_smooth_action(t_act, self.norm_std_policy, self.noise_clip, eps=eps)
            else:
                rng = self._rng.ensure(t_act.device)
                t_act = ops.td3_smooth_action(t_act, self.norm_std_policy, self.noise_clip, rng=rng)
                ops.counter_advance(rng.counter)
            tq1 = self.target_qf1([next_obs, t_act]).reshape(-1)
            tq2 = self.target_qf2([next_obs, t_act]).reshape(-1)
            y, _ = ops.td_target(rewards, terminals, tq1, tq2, None, None, self.discount, sc, info=info[0:1])
        q1_pred = self.qf1([obs, acts])
        q2_pred = self.qf2([obs, acts])
        g1, g2, _ = ops.twin_mse_loss(q1_pred.reshape(-1), q2_pred.reshape(-1), y, sc, info=info[4:6])
        torch.autograd.backward([q1_pred, q2_pred], [g1.reshape(q1_pred.shape), g2.reshape(q2_pred.shape)],
                                inputs=self.opt.segments[1] + self.opt.segments[2])
        self._step(active_mask=0b110)
        if variant == 1:
            new_actions = self.pf(obs)
            q_new = self.qf1([obs, new_actions])            # uses qf1 AFTER its step, like the reference
            info[6:7].copy_((-q_new.detach().mean()).reshape(1))
            seed = torch.full_like(q_new, -1.0 / q_new.numel())
            torch.autograd.backward([q_new], [seed], inputs=self.opt.segments[0])
            self._step(active_mask=0b001)
            self._update_target_networks()
            ops.vec_stats(new_actions.detach().reshape(-1), out=info[10:14])
        if self._explicit_batch is None:
            self._finish_update()

    def _decode_info(self, row, variant):
        info = {'Reward_Mean': float(row[0]), 'Training/qf1_loss': float(row[4]), 'Training/qf2_loss': float(row[5])}
        if variant == 1:
            info['Training/policy_loss'] = float(row[6])
            for i, s in enumerate(_STAT):
                info['new_actions/' + s] = float(row[10 + i])
        return info

    @property
    def networks(self):
        return [self.pf, self.qf1, self.qf2, self.target_pf, self.target_qf1, self.target_qf2]

    @property
    def snapshot_networks(self):
        return [["pf", self.pf], ["qf1", self.qf1], ["qf2", self.qf2]]

    @property
    def target_networks(self):
        return [(self.pf, self.target_pf), (self.qf1, self.target_qf1), (self.qf2, self.target_qf2)]
