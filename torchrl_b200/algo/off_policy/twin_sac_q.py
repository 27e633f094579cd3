"""Twin-Q SAC without a V network (API of /root/reference/torchrl/algo/off_policy/twin_sac_q.py:11-250).

Per update (one captured CUDA graph): row gather -> pf(obs) + sampling/log-prob kernel (autograd) ->
Q1,Q2(obs,a) -> temperature loss + its Adam step (1 launch) -> [no grad] pf(next_obs) sample,
target Q's, TD-target kernel -> twin MSE kernel -> Q1,Q2(obs, a~) -> policy-loss kernel -> two
autograd.backward calls seeded with the kernel gradients -> fused clip+Adam over pf|qf1|qf2 ->
Polyak over the flat target buffer -> log row.
"""
import copy

import numpy as np
import torch
import torch.optim as optim

from ... import ops
from ...flat import FlatAdam, FlatParams
from ...policies import distribution as D
from ..rl_algo import SegmentOptimizer
from .off_rl_algo import OffRLAlgo

_STAT = ("mean", "std", "max", "min")


class TwinSACQ(OffRLAlgo):
    def __init__(self, pf, qf1, qf2, plr, qlr, optimizer_class=optim.Adam, policy_std_reg_weight=1e-3,
                 policy_mean_reg_weight=1e-3, reparameterization=True, automatic_entropy_tuning=True,
                 target_entropy=None, **kwargs):
        super().__init__(**kwargs)
        self.pf, self.qf1, self.qf2 = pf, qf1, qf2
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
        self._target_flat = FlatParams([self.target_qf1, self.target_qf2], device=self.device)
        self.automatic_entropy_tuning = automatic_entropy_tuning
        if self.automatic_entropy_tuning:
            self.target_entropy = target_entropy if target_entropy else \
                -float(np.prod(self.env.action_space.shape).item())
            self.log_alpha = torch.zeros(1, device=self.device)
            self._alpha_state = torch.zeros(3, device=self.device)     # exp_avg, exp_avg_sq, step
        self.policy_std_reg_weight = policy_std_reg_weight
        self.policy_mean_reg_weight = policy_mean_reg_weight
        if not reparameterization:
            raise NotImplementedError
        self.reparameterization = reparameterization
        self.tanh_action = bool(getattr(pf, "tanh_action", True))

    def _target_source(self):
        return self.opt.seg_slice(1, 3)

    # info row layout
    #  0 Reward_Mean | 1 Alpha 2 Alpha_loss | 3 policy_loss(kernel part) 4 qf1_loss 5 qf2_loss
    #  6..9 log_probs mean/std/max/min | 10..13 log_std stats | 14..17 mean stats | 18 std_reg 19 mean_reg
    #  20.. scratch for the 5-float policy info
    def _sample(self, obs, want_grad):
        mean, std, log_std = self.pf(obs)
        mean_c = mean if mean.is_contiguous() else mean.contiguous()
        ls = log_std if log_std.dim() == 1 else (log_std if log_std.is_contiguous() else log_std.contiguous())
        eps = D.draw_reference_noise(tuple(mean_c.shape), mean_c.device) if D.get_noise_mode() == "reference_cpu" \
            else None
        rng = self.pf._rng_state(mean_c.device)
        if want_grad:
            action, logp, _ = D._SampleFn.apply(mean_c, ls, eps, self.tanh_action, True, rng)
        else:
            out = ops.tanh_gaussian_sample(mean_c, ls, eps=eps, tanh_action=self.tanh_action, want_log_prob=True,
                                           rng=rng)
            action, logp = out["action"], out["log_prob"]
        if eps is None:
            ops.counter_advance(rng.counter)
        return action, logp, mean_c, ls

    def _update_body(self, variant):
        ub = self._ub
        batch = self._batch()
        info = ub["info"][0]
        sc = ub["scratch"]
        obs, acts, next_obs = batch["obs"], batch["acts"], batch["next_obs"]
        rewards, terminals = batch["rewards"].reshape(-1), batch["terminals"].reshape(-1)
        B = obs.shape[0]
        acts = acts.reshape(B, -1)
        new_actions, log_probs, mean, log_std = self._sample(obs, True)
        q1_pred = self.qf1([obs, acts])
        q2_pred = self.qf2([obs, acts])
        log_alpha = None
        if self.automatic_entropy_tuning:
            lp_all = self._all_ranks(log_probs.detach().reshape(-1))      # temperature sees every rank's samples
            alpha_sc = sc
            if lp_all.numel() != sc.B:                  # more samples than the per-rank scratch was sized for
                alpha_sc = getattr(self, "_alpha_sc", None)
                if alpha_sc is None or alpha_sc.B != lp_all.numel():
                    alpha_sc = self._alpha_sc = ops.OffPolicyScratch(lp_all.numel(), lp_all.device)
            ops.sac_alpha_step(lp_all, self.target_entropy, self.log_alpha, self._alpha_state, self.plr, alpha_sc,
                               info=info[1:3])
            log_alpha = self.log_alpha
        with torch.no_grad():
            t_actions, t_logp, _, _ = self._sample(next_obs, False)
            tq1 = self.target_qf1([next_obs, t_actions]).reshape(-1)
            tq2 = self.target_qf2([next_obs, t_actions]).reshape(-1)
            y, _ = ops.td_target(rewards, terminals, tq1, tq2, t_logp.reshape(-1), log_alpha, self.discount, sc,
                                 info=info[0:1], fixed_alpha=1.0)
        g1, g2, _ = ops.twin_mse_loss(q1_pred.reshape(-1), q2_pred.reshape(-1), y, sc, info=info[4:6])
        qn1 = self.qf1([obs, new_actions])
        qn2 = self.qf2([obs, new_actions])
        g_lp, g_qn1, g_qn2, _ = ops.sac_policy_loss(log_probs.reshape(-1), qn1.reshape(-1), qn2.reshape(-1),
                                                    log_alpha, sc, info=info[20:25], fixed_alpha=1.0)
        ops.vec_stats(log_std.detach().reshape(-1) if log_std.is_contiguous() else log_std.detach().contiguous().reshape(-1),
                      out=info[10:14])
        ops.vec_stats(mean.detach().reshape(-1), out=info[14:18])
        roots = [log_probs, qn1, qn2]
        seeds = [g_lp.reshape(log_probs.shape), g_qn1.reshape(qn1.shape), g_qn2.reshape(qn2.shape)]
        if self.policy_std_reg_weight or self.policy_mean_reg_weight:
            std_reg = self.policy_std_reg_weight * (log_std ** 2).mean()
            mean_reg = self.policy_mean_reg_weight * (mean ** 2).mean()
            info[18:19].copy_(std_reg.detach().reshape(1))
            info[19:20].copy_(mean_reg.detach().reshape(1))
            roots.append(std_reg + mean_reg)
            seeds.append(torch.ones((), device=obs.device))
        pf_params = self.opt.segments[0]
        torch.autograd.backward(roots, seeds, inputs=pf_params)
        torch.autograd.backward([q1_pred, q2_pred], [g1.reshape(q1_pred.shape), g2.reshape(q2_pred.shape)],
                                inputs=self.opt.segments[1] + self.opt.segments[2])
        self._step()
        self._update_target_networks()
        if self._explicit_batch is None:
            self._finish_update()

    def _decode_info(self, row, variant):
        info = {'Reward_Mean': float(row[0])}
        if self.automatic_entropy_tuning:
            info["Alpha"] = float(row[1])
            info["Alpha_loss"] = float(row[2])
        info['Training/policy_loss'] = float(row[20] + row[18] + row[19])
        info['Training/qf1_loss'] = float(row[4])
        info['Training/qf2_loss'] = float(row[5])
        for i, s in enumerate(_STAT):
            info['log_std/' + s] = float(row[10 + i])
        for i, s in enumerate(_STAT):
            info['log_probs/' + s] = float(row[21 + i])
        for i, s in enumerate(_STAT):
            info['mean/' + s] = float(row[14 + i])
        return info

    @property
    def networks(self):
        return [self.pf, self.qf1, self.qf2, self.target_qf1, self.target_qf2]

    @property
    def snapshot_networks(self):
        return [["pf", self.pf], ["qf1", self.qf1], ["qf2", self.qf2]]

    @property
    def target_networks(self):
        return [(self.qf1, self.target_qf1), (self.qf2, self.target_qf2)]
