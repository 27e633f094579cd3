"""PPO on the device (API of /root/reference/torchrl/algo/on_policy/ppo.py:10-160).

Per minibatch (one captured CUDA graph, replayed opt_epochs * T/b times per epoch):
  row gather of all keys (1 launch) -> advantage stats (1) -> vf forward (torch) -> critic loss
  fwd+bwd kernel (1) -> autograd through vf -> pf forward (torch) -> actor loss fwd+bwd kernel (1)
  -> autograd through pf -> [NCCL all-reduce of the flat gradient] -> grad-norm + clip + Adam
  (2) -> info row store (1) -> device counters (2).
Nothing syncs with the host inside the loop: the 19 logged scalars per update (the reference
does 18 .item() calls, ppo.py:76-91,121-122,141-144) are stored in a device log and fetched
once per epoch.

Exactness notes:
  * old log-probs are computed once per epoch and gathered with the minibatch -- the reference
    recomputes target_pf(obs) every minibatch (ppo.py:54-56) but target_pf is constant within
    an epoch, so the cached values are the same numbers;
  * critic and actor steps of one minibatch are fused into one optimizer launch: the two
    networks share no parameters, so the result equals the reference's critic-then-actor order;
  * minibatch row order comes from np.random.permutation on the host (bit-exact indexing).
"""
import copy

import numpy as np
import torch

from ... import ops
from ...flat import FlatParams
from ...networks import fused
from .. import utils as atu
from .a2c import A2C

_INFO_KEYS_ACTOR = ['Training/policy_loss', 'logprob/mean', 'logprob/std', 'logprob/max', 'logprob/min',
                    'ratio/max', 'ratio/min', 'log_std/mean', 'log_std/std', 'log_std/max', 'log_std/min']
_ADV_KEYS = ['advs/mean', 'advs/std', 'advs/max', 'advs/min']


class PPO(A2C):
    def __init__(self, pf, clip_para=0.2, opt_epochs=10, clipped_value_loss=False, **kwargs):
        self.target_pf = copy.deepcopy(pf)
        super().__init__(pf=pf, **kwargs)
        self.clip_para = clip_para
        self.opt_epochs = opt_epochs
        self.clipped_value_loss = clipped_value_loss
        self.sample_key = ["obs", "acts", "advs", "estimate_returns", "values"]
        self._target_flat = FlatParams([self.target_pf], device=self.device)
        self._mb_graph = None
        self._mb_eager_runs = 0
        self._mb_state = None
        self.tanh_action = bool(getattr(pf, "tanh_action", False))

    # ------------------------------------------------------------------ helpers
    def _policy_outputs(self, pf, obs):
        mean, _, log_std = pf(obs)
        if not mean.is_contiguous():
            mean = mean.contiguous()
        if not log_std.is_contiguous():
            log_std = log_std.contiguous()
        return mean, log_std

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

    def _mb_setup(self):
        rb = self.replay_buffer
        N = rb.env_nums
        assert self.batch_size % N == 0, "batch size should be dividable by env_nums"
        b = self.batch_size // N
        T = rb._max_replay_buffer_size
        assert T % b == 0, "rows per minibatch must divide the buffer rows"
        n_mb = T // b
        U = self.opt_epochs * n_mb
        dev = self.device
        a = rb._acts.shape[-1]
        st = {
            "b": b, "n_mb": n_mb, "U": U, "B": b * N,
            # every pass' row order is uploaded up-front: (opt_epochs, T) indices, minibatch u of the
            # epoch reads perm[u*b : (u+1)*b] -- `upd` is both the gather position and the log row
            "perm": torch.zeros(self.opt_epochs * T, dtype=torch.int64, device=dev),
            "perm_host": torch.zeros(self.opt_epochs * T, dtype=torch.int64).pin_memory(),
            "upd": torch.zeros(1, dtype=torch.int32, device=dev),
            "info": torch.zeros(1, 32, dtype=torch.float32, device=dev),
            "log32": torch.zeros(U, 32, dtype=torch.float32, device=dev),
            "log64": torch.zeros(U, self.opt.sumsq3.numel(), dtype=torch.float64, device=dev),
            "scratch": ops.LossScratch(b * N, a, dev),
            # advantage statistics of all U minibatches, computed once per epoch (mean, std, max, min per row)
            "adv_table": torch.zeros(U, 4, dtype=torch.float32, device=dev),
            "keys": ["obs", "acts", "advs", "estimate_returns", "values", "old_logp"],
        }
        st["log_plan"] = ops.RowCopyPlan([st["info"], self.opt.sumsq3.view(1, -1)], [st["log32"], st["log64"]],
                                         [32 * 4, self.opt.sumsq3.numel() * 8])
        self._mb_state = st
        return st

    def _mb_body(self):
        """One minibatch update reading its row indices at device position `upd`.  Layer gradients go
        straight into the flat gradient buffer (networks.fused.direct_grad): every parameter gets exactly
        one contribution per minibatch and the fused Adam step left the buffer zeroed."""
        with fused.direct_grad():
            self._mb_body_inner()

    def _mb_body_inner(self):
        st, rb = self._mb_state, self.replay_buffer
        batch = rb.gather_rows(st["perm"], st["keys"], pos_ptr=st["upd"], rows=st["b"])
        info = st["info"][0]
        advs = batch["advs"].reshape(-1)
        # critic
        v = self.vf(batch["obs"])
        g_v, _ = ops.ppo_critic_loss(v.reshape(-1), batch["estimate_returns"].reshape(-1),
                                     batch["values"].reshape(-1), self.clipped_value_loss, self.clip_para,
                                     st["scratch"], info=info[16:17])
        torch.autograd.backward([v], [g_v.reshape(v.shape)])
        # actor
        mean, log_std = self._policy_outputs(self.pf, batch["obs"])
        g_mean, g_ls, _ = ops.ppo_actor_loss(mean, log_std, batch["acts"].reshape(mean.shape[0], -1),
                                             batch["old_logp"].reshape(-1), advs, st["adv_table"], self.clip_para,
                                             self.entropy_coeff, self.tanh_action, st["scratch"], info=info[0:16],
                                             stats_pos=st["upd"])
        torch.autograd.backward([mean, log_std], [g_mean, g_ls])
        # gradient exchange (multi-GPU) + clip + Adam + zero_grad
        scale, fused_norm = 1.0, False
        if self.dist is not None:
            scale, fused_norm = self.dist.reduce_grads(self.opt)
        self.opt.step(grad_scale=scale, reduced=fused_norm)
        # per-update log row, device counters
        ops.ring_write(st["log_plan"], st["upd"])
        ops.counter_advance(None, st["upd"], st["U"])

    def _run_minibatch(self):
        if not self.use_cuda_graph:
            self._mb_body()
        elif self._mb_graph is not None:
            self._mb_graph.replay()
        elif self._mb_eager_runs < 3:
            self._mb_eager_runs += 1
            self._mb_body()
        else:
            g = ops.CapturedGraph(self._mb_body)
            self._mb_graph = g
            g.replay()
        self.training_update_num += 1

    def _epoch_adv_stats(self):
        """ppo.py:141-147 for ALL minibatches of the epoch at once: which time rows form minibatch u is known as
        soon as the permutations are uploaded, so one launch reduces every minibatch's advantages to raw moments,
        (data parallel) ONE exchange gathers the ranks' moments, one launch turns them into the (U,4) table the actor
        loss indexes with the device counter.  Per minibatch this removes a reduction launch and, with several
        ranks, an all-gather."""
        st, rb = self._mb_state, self.replay_buffer
        U, b = st["U"], st["b"]
        dp = self.dist is not None and self.dist.active
        W = self.dist.world_size if dp else 1
        if "mom_all" not in st:
            st["mom_all"] = torch.zeros(W, U, 4, dtype=torch.float64, device=self.device)
            if dp and self.dist.peer is not None:
                st["mom"] = self.dist.peer.region("adv_moments", 32 * U, torch.float64)[0][:4 * U].view(U, 4)
            else:
                st["mom"] = torch.zeros(U, 4, dtype=torch.float64, device=self.device) if dp else st["mom_all"][0]
        advs = rb._advs.reshape(rb._advs.shape[0], -1)
        ops.row_group_moments(advs, st["perm"], U, b, out=st["mom"])
        if dp:
            if self.dist.peer is not None:
                self.dist.peer.all_reduce_f64("adv_moments", 4 * U, st["mom_all"], gather=True)
            else:
                import torch.distributed as tdist
                tdist.all_gather_into_tensor(st["mom_all"].view(-1), st["mom"].view(-1))
        ops.group_stats_from_moments(st["mom_all"], W, U, float(b * rb.env_nums * W), out=st["adv_table"])

    def _flush_infos(self, n_updates):
        """One D2H copy of the epoch's per-update scalars -> list of the reference's info dicts."""
        st = self._mb_state
        log32 = st["log32"][:n_updates].cpu().numpy()
        log32[:, 20:24] = st["adv_table"][:n_updates].cpu().numpy()
        log64 = st["log64"][:n_updates].cpu().numpy()
        infos = []
        for u in range(n_updates):
            row = log32[u]
            info = {}
            for i, k in enumerate(_ADV_KEYS):
                info[k] = float(row[20 + i])
            info['Training/vf_loss'] = float(row[16])
            # the flat gradient holds the SUM over ranks; the averaged gradient's norm is what one process sees
            gs = 1.0 / self.dist.world_size if (self.dist is not None and self.dist.active) else 1.0
            info['grad_norm/vf'] = float(np.sqrt(log64[u][1])) * gs
            for i, k in enumerate(_INFO_KEYS_ACTOR):
                info[k] = float(row[i])
            info['grad_norm/pf'] = float(np.sqrt(log64[u][0])) * gs
            infos.append(info)
        return infos

    # ------------------------------------------------------------------ reference API
    @fused.presplit_scope
    def update_per_epoch(self, flush_infos=True):
        """ppo.py:27-39: GAE, linear LR decay, target <- pf, opt_epochs passes of minibatches.
        flush_infos=False skips the end-of-epoch read-back of the logged scalars (benchmarking the
        device path alone)."""
        self.process_epoch_samples()
        atu.update_linear_schedule(self.pf_optimizer, self.current_epoch, self.num_epochs, self.plr)
        atu.update_linear_schedule(self.vf_optimizer, self.current_epoch, self.num_epochs, self.vlr)
        self._target_flat.copy_from(self.opt.seg_slice(0))       # copy_model_params_from_to(pf, target_pf)
        self._cache_old_logp()
        st = self._mb_state or self._mb_setup()
        st["upd"].zero_()
        T = self.replay_buffer._max_replay_buffer_size
        # the reference draws one np.random.permutation per pass, nothing else touches np.random
        # in between, so drawing all passes up-front consumes the global RNG identically
        for e in range(self.opt_epochs):
            order = self.replay_buffer.epoch_order(self.shuffle)
            st["perm_host"][e * T:(e + 1) * T].copy_(torch.from_numpy(np.ascontiguousarray(order, dtype=np.int64)))
        st["perm"].copy_(st["perm_host"], non_blocking=True)
        self._epoch_adv_stats()
        n = st["U"]
        for _ in range(n):
            self._run_minibatch()
        if not flush_infos:
            return
        self._last_infos = self._flush_infos(n)
        if self.logger is not None:
            for info in self._last_infos:
                self.logger.add_update_info(info)

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
        torch.autograd.backward([v], [g_v.reshape(v.shape)])
        mean, log_std = self._policy_outputs(self.pf, obs)
        g_mean, g_ls, _ = ops.ppo_actor_loss(mean, log_std, acts, old_logp, advs.reshape(-1), adv_stats,
                                             self.clip_para, self.entropy_coeff, self.tanh_action, scratch,
                                             info=info32[0:16])
        torch.autograd.backward([mean, log_std], [g_mean, g_ls])
        scale, fused_norm = 1.0, False
        if self.dist is not None:
            scale, fused_norm = self.dist.reduce_grads(self.opt)
        self.opt.step(grad_scale=scale, reduced=fused_norm)
        row = info32.cpu().numpy()
        norms = self.opt.grad_norms().cpu().numpy() * scale
        info = {k: float(row[20 + i]) for i, k in enumerate(_ADV_KEYS)}
        info['Training/vf_loss'] = float(row[16])
        info['grad_norm/vf'] = float(norms[1])
        for i, k in enumerate(_INFO_KEYS_ACTOR):
            info[k] = float(row[i])
        info['grad_norm/pf'] = float(norms[0])
        return info

    @property
    def networks(self):
        return [self.pf, self.vf, self.target_pf]
