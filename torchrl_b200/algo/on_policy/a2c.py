"""Advantage actor-critic on the device (API of /root/reference/torchrl/algo/on_policy/a2c.py:8-112) and the
minibatch machinery PPO shares with it.

Per minibatch (one captured CUDA graph, replayed T/b times per pass):
  row gather of all keys (1 launch) -> vf forward -> critic loss fwd+bwd kernel -> autograd through vf -> value
  statistics (1) -> pf forward -> actor loss fwd+bwd kernel (policy-gradient mode: L = -mean(logp * adv_norm) -
  c_ent * mean(ent), a2c.py:66-70) -> autograd through pf -> [gradient exchange] -> grad-norm + clip(0.5) + Adam
  (eps 1e-5) for both networks in one step (a2c.py:29-39, 72-80: two optimizers over disjoint parameters) -> info row.
The advantage statistics of every minibatch of the epoch are computed once up-front (`_epoch_adv_stats`).  Nothing
syncs with the host inside the loop: the reference's 12 `.item()` calls per update (a2c.py:82-98) become one device
log fetched per epoch.  The constructor re-homes pf and vf into one flat buffer (flat.FlatAdam).
"""
import os

import numpy as np
import torch
import torch.optim as optim

from ... import ops
from ...flat import FlatAdam
from ...networks import fused
from ..rl_algo import SegmentOptimizer
from .on_rl_algo import OnRLAlgo

_HALF_LOG_2PI = 0.5 * float(np.log(2.0 * np.pi))
_ADV_KEYS = ['advs/mean', 'advs/std', 'advs/max', 'advs/min']


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
        # segment 0 = policy, segment 1 = value net (+ whatever a subclass optimises besides, e.g. V-MPO's duals)
        extra = self._extra_opt_segments()
        self.opt = FlatAdam([self.pf, self.vf] + [e[0] for e in extra], lrs=[plr, vlr] + [e[1] for e in extra], eps=1e-5,
                            max_norms=[0.5, 0.5] + [e[2] for e in extra], device=self.device, dist=self.dist)
        self.pf_optimizer = SegmentOptimizer(self.opt, 0)
        self.vf_optimizer = SegmentOptimizer(self.opt, 1)
        self.entropy_coeff = entropy_coeff
        self.vf_criterion = torch.nn.MSELoss()
        self.sample_key = ["obs", "acts", "advs", "estimate_returns"]
        self.tanh_action = bool(getattr(pf, "tanh_action", False))
        self._mb_graph = None
        self._mb_eager_runs = 0
        self._mb_state = None
        self._last_infos = []
        self.overlap_nets = os.environ.get("TORCHRL_B200_OVERLAP_NETS", "1") == "1"
        self._side_stream = torch.cuda.Stream(device=self.device)

    # ------------------------------------------------------------------ what subclasses specialise
    def _extra_opt_segments(self):
        """[(parameter list, lr, max grad norm or 0)] optimised by the same fused step besides pf and vf."""
        return []

    def _step_mask(self):
        """Bit mask of the optimizer segments that step in the minibatch loop (None: all)."""
        return None

    def _passes(self):
        """Optimisation passes over the epoch's rollout (a2c: one; ppo: opt_epochs)."""
        return 1

    def _gather_keys(self):
        return ["obs", "acts", "advs", "estimate_returns"]

    def _critic_step(self, batch, info):
        v = self.vf(batch["obs"])
        g_v, _ = ops.ppo_critic_loss(v.reshape(-1), batch["estimate_returns"].reshape(-1), None, False, 0.0,
                                     self._mb_state["scratch"], info=info[16:17])
        with fused.backward_fork():
            torch.autograd.backward([v], [g_v.reshape(v.shape)])
        ops.vec_stats(v.detach().reshape(-1), out=info[24:28])          # v_pred/* (a2c.py:85-88)

    def _actor_step(self, batch, info):
        st = self._mb_state
        mean, raw_ls, clamp, g_ls = self._raw_policy_outputs(self.pf, batch["obs"])
        g_mean, _, _ = ops.ppo_actor_loss(mean, raw_ls.detach(), batch["acts"].reshape(mean.shape[0], -1), None,
                                          batch["advs"].reshape(-1), st["adv_table"], 0.0, self.entropy_coeff,
                                          self.tanh_action, st["scratch"], g_log_std=g_ls, info=info[0:16],
                                          stats_pos=st["upd"], ls_clamp=clamp)
        with fused.backward_fork():
            torch.autograd.backward([mean], [g_mean])
        # std/* (a2c.py:90-94) derive from the clamped log-std at flush
        torch.clamp(raw_ls.detach(), clamp[0], clamp[1], out=info[28:28 + raw_ls.numel()])

    def _pre_update(self):
        """Host-side work of an epoch before the minibatch loop (schedules, target copies)."""

    def _decode_info(self, row, norms, gs):
        a = self.replay_buffer._acts.shape[-1]
        ls = row[28:28 + a].astype(np.float64)
        sd = np.exp(ls)
        B = self._mb_state["B"]
        m = sd.mean()
        var = B * ((sd - m) ** 2).sum() / (B * a - 1.0)              # torch.std() of the (B, a) expanded tensor
        return {'Training/policy_loss': float(row[0]), 'Training/vf_loss': float(row[16]),
                'v_pred/mean': float(row[24]), 'v_pred/std': float(row[25]), 'v_pred/max': float(row[26]),
                'v_pred/min': float(row[27]), 'std/mean': float(m), 'std/std': float(np.sqrt(var)),
                'std/max': float(sd.max()), 'std/min': float(sd.min()), 'ent': float(row[11]),
                'log_prob': float(row[1])}

    # ------------------------------------------------------------------ helpers
    def _policy_outputs(self, pf, obs):
        mean, _, log_std = pf(obs)
        if not mean.is_contiguous():
            mean = mean.contiguous()
        if not log_std.is_contiguous():
            log_std = log_std.contiguous()
        return mean, log_std

    def _raw_policy_outputs(self, pf, obs):
        """Device minibatch path (a shared log-std PARAMETER, _device_path_ok): the mean, the raw parameter, its clamp
        range and the slice of the flat gradient buffer that belongs to it.  The loss kernel applies the policy's
        torch.clamp itself and writes the parameter's gradient in place, which removes the clamp / exp / clamp-backward
        / accumulate launches on six-element tensors from every minibatch."""
        from ...policies.continuous_policy import LOG_SIG_MIN, LOG_SIG_MAX
        mean = pf.mean_net(obs)
        if not mean.is_contiguous():
            mean = mean.contiguous()
        return mean, pf.logstd, (LOG_SIG_MIN, LOG_SIG_MAX), pf.logstd.grad

    def _device_path_ok(self):
        """The fused minibatch loop needs a Gaussian policy with a shared log-std vector (GuassianContPolicyBasicBias)
        over a device rollout buffer; anything else takes the eager `update(batch)` route."""
        rb = self.replay_buffer
        return hasattr(self.pf, "logstd") and rb is not None and hasattr(rb, "gather_rows") and hasattr(rb, "_rewards")

    def _mb_setup(self):
        rb = self.replay_buffer
        N = rb.env_nums
        assert self.batch_size % N == 0, "batch size should be dividable by env_nums"
        b = self.batch_size // N
        T = rb._max_replay_buffer_size
        assert T % b == 0, "rows per minibatch must divide the buffer rows"
        n_mb = T // b
        passes = self._passes()
        U = passes * n_mb
        dev = self.device
        a = rb._acts.shape[-1]
        st = {
            "b": b, "n_mb": n_mb, "U": U, "B": b * N, "passes": passes,
            # every pass' row order is uploaded up-front: (passes, T) indices, minibatch u of the epoch reads
            # perm[u*b : (u+1)*b] -- `upd` is the gather position, the statistics row and the log row
            "perm": torch.zeros(passes * T, dtype=torch.int64, device=dev),
            "perm_host": torch.zeros(passes * T, dtype=torch.int64).pin_memory(),
            "upd": torch.zeros(1, dtype=torch.int32, device=dev),
            "info": torch.zeros(1, 64, dtype=torch.float32, device=dev),
            "log_ticket": torch.zeros(1, dtype=torch.int32, device=dev),
            "log32": torch.zeros(U, 64, dtype=torch.float32, device=dev),
            "log64": torch.zeros(U, self.opt.sumsq3.numel(), dtype=torch.float64, device=dev),
            "scratch": ops.LossScratch(b * N, a, dev),
            # advantage statistics of all U minibatches, computed once per epoch (mean, std, max, min per row)
            "adv_table": torch.zeros(U, 4, dtype=torch.float32, device=dev),
            "keys": self._gather_keys(),
        }
        st["log_plan"] = ops.RowCopyPlan([st["info"], self.opt.sumsq3.view(1, -1)], [st["log32"], st["log64"]],
                                         [64 * 4, self.opt.sumsq3.numel() * 8])
        self._mb_state = st
        return st

    def _mb_body(self):
        """One minibatch update reading its row indices at device position `upd`.  Layer gradients go straight into
        the flat gradient buffer (networks.fused.direct_grad): every parameter gets exactly one contribution per
        minibatch and the optimizer step left the buffer zeroed."""
        with fused.direct_grad(), fused.deferred_reduces():
            st, rb = self._mb_state, self.replay_buffer
            batch = rb.gather_rows(st["perm"], st["keys"], pos_ptr=st["upd"], rows=st["b"])
            info = st["info"][0]
            if self.overlap_nets:
                # the critic and the actor branch share nothing but their (read-only) inputs: run them on two streams
                # -- under capture this becomes two parallel branches of the graph -- so that their many
                # latency-bound launches overlap instead of queueing behind one another
                main = torch.cuda.current_stream(self.device)
                side = self._side_stream
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self._critic_step(batch, info)
                self._actor_step(batch, info)
                main.wait_stream(side)
            else:
                self._critic_step(batch, info)
                self._actor_step(batch, info)
            fused.flush_reduces()              # the slab sums of both networks' skinny gradients, one launch
            scale, fused_norm = 1.0, False
            if self.dist is not None:
                scale, fused_norm = self.dist.reduce_grads(self.opt, self._step_mask())   # exchange + norms, one kernel
            self.opt.step(active_mask=self._step_mask(), grad_scale=scale, reduced=fused_norm)
            ops.ring_write_advance(st["log_plan"], st["upd"], st["U"], st["log_ticket"])   # log row, then upd += 1

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
        """a2c.py:62 / ppo.py:141-147 for ALL minibatches of the epoch at once: which time rows form minibatch u is
        known as soon as the permutations are uploaded, so one launch reduces every minibatch's advantages to raw
        moments, (data parallel) ONE exchange gathers the ranks' moments, one launch turns them into the (U,4) table
        the actor loss indexes with the device counter.  Per minibatch this removes a reduction launch and, with
        several ranks, an all-gather."""
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
        # the flat gradient holds the SUM over ranks; the averaged gradient's norm is what one process sees
        gs = 1.0 / self.dist.world_size if (self.dist is not None and self.dist.active) else 1.0
        return [self._decode_info(log32[u], np.sqrt(log64[u][:self.opt.nseg]) * gs, gs) for u in range(n_updates)]

    # ------------------------------------------------------------------ reference API
    @fused.presplit_scope
    def update_per_epoch(self, flush_infos=True):
        """on_rl_algo.py:35-42 (a2c) / ppo.py:27-39: advantages, then `passes` sweeps of row-order minibatches.
        flush_infos=False skips the end-of-epoch read-back of the logged scalars (device-only benchmarking)."""
        if not self._device_path_ok():
            return super().update_per_epoch()
        self.process_epoch_samples()
        self._pre_update()
        st = self._mb_state or self._mb_setup()
        st["upd"].zero_()
        T = self.replay_buffer._max_replay_buffer_size
        # the reference draws one np.random.permutation per pass and nothing else touches np.random in between, so
        # drawing all passes up-front consumes the global RNG identically
        for e in range(st["passes"]):
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

    def _minibatch(self, batch, keys):
        return [torch.as_tensor(np.asarray(batch[k]) if not torch.is_tensor(batch[k]) else batch[k],
                                dtype=torch.float32, device=self.device).contiguous() for k in keys]

    @staticmethod
    def _four_stats(prefix, t):
        return {prefix + '/mean': t.mean().item(), prefix + '/std': t.std().item(),
                prefix + '/max': t.max().item(), prefix + '/min': t.min().item()}

    @fused.presplit_scope
    def update(self, batch):
        """One A2C minibatch update on an explicit batch (a2c.py:45-112), eagerly, through the same loss kernels and
        the fused optimizer step; returns the reference's info dict (this entry point syncs; the epoch loop does not
        use it)."""
        self.training_update_num += 1
        obs, acts, advs, est_rets = self._minibatch(batch, ('obs', 'acts', 'advs', 'estimate_returns'))
        B = obs.shape[0]
        acts = acts.reshape(B, -1)
        scratch = ops.LossScratch(B, acts.shape[1], self.device)
        info32 = torch.zeros(32, dtype=torch.float32, device=self.device)
        adv_stats = ops.vec_stats(advs.reshape(-1), out=info32[20:24])
        values = self.vf(obs)
        g_v, _ = ops.ppo_critic_loss(values.reshape(-1), est_rets.reshape(-1), None, False, 0.0, scratch, info=info32[16:17])
        torch.autograd.backward([values], [g_v.reshape(values.shape)])
        mean, std, log_std = self.pf(obs)
        mean = mean if mean.is_contiguous() else mean.contiguous()
        ls = log_std if log_std.is_contiguous() else log_std.contiguous()
        if ls.dim() > 1 and ls.shape != mean.shape:
            ls = ls.expand_as(mean).contiguous()
        g_mean, g_ls, _ = ops.ppo_actor_loss(mean, ls, acts, None, advs.reshape(-1), adv_stats, 0.0, self.entropy_coeff,
                                             self.tanh_action, scratch, info=info32[0:16])
        torch.autograd.backward([mean, ls], [g_mean, g_ls])
        scale, fused_norm = 1.0, False
        if self.dist is not None:
            scale, fused_norm = self.dist.reduce_grads(self.opt)
        self.opt.step(grad_scale=scale, reduced=fused_norm)
        row = info32.cpu().numpy()
        info = {'Training/policy_loss': float(row[0]), 'Training/vf_loss': float(row[16])}
        info.update(self._four_stats('v_pred', values.detach()))
        info.update(self._four_stats('std', std.detach().expand_as(mean)))
        info['ent'] = float(row[11])
        info['log_prob'] = float(row[1])
        return info

    @property
    def snapshot_networks(self):
        return [("pf", self.pf), ("vf", self.vf)]
