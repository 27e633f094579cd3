"""PPO on the device (API of /root/reference/torchrl/algo/on_policy/ppo.py:10-160).

The per-minibatch loop (gather -> critic loss -> actor loss -> [gradient exchange] -> clip + Adam -> log row, one
captured CUDA graph replayed opt_epochs * T/b times per epoch, no host sync inside) lives in a2c.A2C; PPO adds the
clipped surrogate (the actor kernel's ratio mode), the optional clipped value loss, the linear LR decay and the
target policy.

Exactness notes:
  * old log-probs are computed once per epoch and gathered with the minibatch -- the reference recomputes
    target_pf(obs) every minibatch (ppo.py:54-56) but target_pf is constant within an epoch, so the cached values are
    the same numbers;
  * critic and actor steps of one minibatch are fused into one optimizer launch: the two networks share no parameters,
    so the result equals the reference's critic-then-actor order;
  * minibatch row order comes from np.random.permutation on the host (bit-exact indexing).
"""
import copy

import numpy as np
import torch

from ... import ops
from ...flat import FlatParams
from ...networks import fused
from .. import utils as atu
from .a2c import A2C, _ADV_KEYS

_INFO_KEYS_ACTOR = ['Training/policy_loss', 'logprob/mean', 'logprob/std', 'logprob/max', 'logprob/min',
                    'ratio/max', 'ratio/min', 'log_std/mean', 'log_std/std', 'log_std/max', 'log_std/min']


class PPO(A2C):
    def __init__(self, pf, clip_para=0.2, opt_epochs=10, clipped_value_loss=False, **kwargs):
        self.target_pf = copy.deepcopy(pf)
        super().__init__(pf=pf, **kwargs)
        self.clip_para = clip_para
        self.opt_epochs = opt_epochs
        self.clipped_value_loss = clipped_value_loss
        self.sample_key = ["obs", "acts", "advs", "estimate_returns", "values"]
        self._target_flat = FlatParams([self.target_pf], device=self.device)

    # ------------------------------------------------------------------ specialisation of the minibatch loop
    def _passes(self):
        return self.opt_epochs

    def _gather_keys(self):
        return ["obs", "acts", "advs", "estimate_returns", "values", "old_logp"]

    def _device_path_ok(self):
        return True

    def _critic_step(self, batch, info):
        v = self.vf(batch["obs"])
        g_v, _ = ops.ppo_critic_loss(v.reshape(-1), batch["estimate_returns"].reshape(-1),
                                     batch["values"].reshape(-1), self.clipped_value_loss, self.clip_para,
                                     self._mb_state["scratch"], info=info[16:17])
        with fused.backward_fork():
            torch.autograd.backward([v], [g_v.reshape(v.shape)])

    def _actor_step(self, batch, info):
        st = self._mb_state
        mean, raw_ls, clamp, g_ls = self._raw_policy_outputs(self.pf, batch["obs"])
        g_mean, _, _ = ops.ppo_actor_loss(mean, raw_ls.detach(), batch["acts"].reshape(mean.shape[0], -1),
                                          batch["old_logp"].reshape(-1), batch["advs"].reshape(-1), st["adv_table"],
                                          self.clip_para, self.entropy_coeff, self.tanh_action, st["scratch"],
                                          g_log_std=g_ls, info=info[0:16], stats_pos=st["upd"], ls_clamp=clamp)
        with fused.backward_fork():
            torch.autograd.backward([mean], [g_mean])

    def _cache_old_logp(self):
        """log pi_old(a|s) for every stored transition, once per epoch (see module docstring)."""
        rb = self.replay_buffer
        if not hasattr(rb, "_old_logp"):
            rb.allocate("old_logp", tuple(rb._rewards.shape[1:]))
        T, N = rb._obs.shape[0], rb._obs.shape[1]
        rows = max(1, (1 << 16) // N)
        with torch.no_grad():
            for r0 in range(0, T, rows):
                r1 = min(T, r0 + rows)
                obs = rb._obs[r0:r1].reshape(-1, rb._obs.shape[-1])
                acts = rb._acts[r0:r1].reshape(obs.shape[0], -1)
                mean, log_std = self._policy_outputs(self.pf, obs)
                ops.gaussian_log_prob(mean, log_std, acts, self.tanh_action,
                                      out=rb._old_logp[r0:r1].reshape(-1))

    def _pre_update(self):
        """ppo.py:29-34: linear LR decay, target <- pf; then the epoch's old log-probs."""
        atu.update_linear_schedule(self.pf_optimizer, self.current_epoch, self.num_epochs, self.plr)
        atu.update_linear_schedule(self.vf_optimizer, self.current_epoch, self.num_epochs, self.vlr)
        self._target_flat.copy_from(self.opt.seg_slice(0))       # copy_model_params_from_to(pf, target_pf)
        self._cache_old_logp()

    def _decode_info(self, row, norms, gs):
        info = {}
        for i, k in enumerate(_ADV_KEYS):
            info[k] = float(row[20 + i])
        info['Training/vf_loss'] = float(row[16])
        info['grad_norm/vf'] = float(norms[1])
        for i, k in enumerate(_INFO_KEYS_ACTOR):
            info[k] = float(row[i])
        info['grad_norm/pf'] = float(norms[0])
        return info

    # ------------------------------------------------------------------ reference API
    @fused.presplit_scope
    def update(self, batch):
        """Eager single-minibatch update with the reference's signature (ppo.py:124-152): `batch`
        holds (B,D) arrays for obs, acts, advs, estimate_returns, values; returns the info dict of
        Python floats (this entry point syncs; the epoch loop does not use it)."""
        self.training_update_num += 1
        dev = self.device
        f = lambda k: torch.as_tensor(np.asarray(batch[k]) if not torch.is_tensor(batch[k]) else batch[k],
                                      dtype=torch.float32, device=dev).contiguous()
        obs, acts, advs, rets, old_v = f('obs'), f('acts'), f('advs'), f('estimate_returns'), f('values')
        B = obs.shape[0]
        acts = acts.reshape(B, -1)
        scratch = ops.LossScratch(B, acts.shape[1], dev)
        info32 = torch.zeros(32, dtype=torch.float32, device=dev)
        adv_stats = ops.vec_stats(advs.reshape(-1), out=info32[20:24])
        if 'old_logp' in batch:
            old_logp = f('old_logp').reshape(-1)
        else:
            with torch.no_grad():
                tmean, tls = self._policy_outputs(self.target_pf, obs)
                old_logp = ops.gaussian_log_prob(tmean, tls, acts, self.tanh_action)
        v = self.vf(obs)
        g_v, _ = ops.ppo_critic_loss(v.reshape(-1), rets.reshape(-1), old_v.reshape(-1), self.clipped_value_loss,
                                     self.clip_para, scratch, info=info32[16:17])
        with fused.backward_fork():
            torch.autograd.backward([v], [g_v.reshape(v.shape)])
        mean, log_std = self._policy_outputs(self.pf, obs)
        g_mean, g_ls, _ = ops.ppo_actor_loss(mean, log_std, acts, old_logp, advs.reshape(-1), adv_stats,
                                             self.clip_para, self.entropy_coeff, self.tanh_action, scratch,
                                             info=info32[0:16])
        with fused.backward_fork():
            torch.autograd.backward([mean, log_std], [g_mean, g_ls])
        scale, fused_norm = 1.0, False
        if self.dist is not None:
            scale, fused_norm = self.dist.reduce_grads(self.opt)
        self.opt.step(grad_scale=scale, reduced=fused_norm)
        row = info32.cpu().numpy()
        norms = self.opt.grad_norms().cpu().numpy() * scale
        return self._decode_info(row, norms, scale)

    @property
    def networks(self):
        return [self.pf, self.vf, self.target_pf]
