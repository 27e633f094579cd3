"""Off-policy base: uniform row sampling + opt_times updates per epoch, target networks
(API of /root/reference/torchrl/algo/off_policy/off_rl_algo.py:8-84).

Device form of `update_per_epoch`: the row indices of ALL opt_times minibatches are drawn up-front
with the reference's own np.random.randint calls (same global RNG stream, bit-exact indices),
uploaded once, and one captured CUDA graph of {gather, update} is replayed opt_times times reading
its position from a device counter.  Logged scalars accumulate in a device log and are fetched
once per epoch (the reference syncs ~20 times per update).
"""
import time

import numpy as np
import torch

from ... import ops
from ...networks import fused
from ..rl_algo import RLAlgo


class OffRLAlgo(RLAlgo):
    INFO_SLOTS = 64

    def __init__(self, pretrain_epochs=0, min_pool=0, target_hard_update_period=1000, use_soft_update=True,
                 tau=0.001, opt_times=1, **kwargs):
        super().__init__(**kwargs)
        self.pretrain_epochs = pretrain_epochs
        self.target_hard_update_period = target_hard_update_period
        self.use_soft_update = use_soft_update
        self.tau = tau
        self.opt_times = opt_times
        self.min_pool = min_pool
        self.sample_key = ["obs", "next_obs", "acts", "rewards", "terminals"]
        self._ub = None              # update-loop state (buffers, graphs)
        self._graphs = {}
        self._eager_runs = {}
        self._last_infos = []

    # ------------------------------------------------------------------ device update loop
    def _ub_setup(self):
        rb = self.replay_buffer
        N = rb.env_nums
        assert self.batch_size % N == 0, "batch size should be dividable by env_nums"
        b = self.batch_size // N
        dev = self.device
        U = max(int(self.opt_times), 1)
        self._ub = {
            "b": b, "B": b * N, "U": U,
            "idx": torch.zeros(U * b, dtype=torch.int64, device=dev),
            "idx_host": torch.zeros(U * b, dtype=torch.int64).pin_memory(),
            "upd": torch.zeros(1, dtype=torch.int32, device=dev),
            "log_ticket": torch.zeros(1, dtype=torch.int32, device=dev),
            "info": torch.zeros(1, self.INFO_SLOTS, dtype=torch.float32, device=dev),
            "log32": torch.zeros(U, self.INFO_SLOTS, dtype=torch.float32, device=dev),
            "scratch": ops.OffPolicyScratch(b * N, dev),
        }
        self._ub["log_plan"] = ops.RowCopyPlan([self._ub["info"]], [self._ub["log32"]], [self.INFO_SLOTS * 4])
        return self._ub

    def _gather(self):
        ub = self._ub
        return self.replay_buffer.gather_rows(ub["idx"], self.sample_key, pos_ptr=ub["upd"], rows=ub["b"])

    def _finish_update(self):
        ub = self._ub
        ops.ring_write_advance(ub["log_plan"], ub["upd"], ub["U"], ub["log_ticket"])     # log row, then upd += 1

    def _variant(self):
        """Key of the update-graph variant for the current update (e.g. TD3's delayed actor step)."""
        return 0

    # ------------------------------------------------------------------ data parallel
    @property
    def _dp(self):
        return self.dist is not None and self.dist.active

    def _step(self, active_mask=None):
        """Optimizer step of the flat buffer.  Data parallel (ring sharded by env, every rank draws the same row
        indices from the same host seed): the flat gradient is summed over ranks first and scaled by 1/G inside
        the Adam kernel, before clipping -- what a single process over all envs would apply.  NOT yet exercised on
        more than one GPU (the single-process path is unchanged)."""
        scale, fused_norm = self.dist.reduce_grads(self.opt, active_mask) if self._dp else (1.0, False)
        self.opt.step(active_mask=active_mask, grad_scale=scale, reduced=fused_norm)

    def _all_ranks(self, vec):
        """Concatenation of a per-rank (B,) vector over ranks (rank order), identical on every rank."""
        if not self._dp:
            return vec
        import torch.distributed as tdist
        out = torch.empty(vec.numel() * self.dist.world_size, dtype=vec.dtype, device=vec.device)
        tdist.all_gather_into_tensor(out, vec.contiguous())
        return out

    def _update_body(self, variant):
        raise NotImplementedError

    def _decode_info(self, row, variant):
        raise NotImplementedError

    def _run_update(self):
        self.training_update_num += 1
        v = self._variant()
        if not self.use_cuda_graph:
            self._update_body(v)
        elif v in self._graphs:
            self._graphs[v].replay()
        elif self._eager_runs.get(v, 0) < 3:
            self._eager_runs[v] = self._eager_runs.get(v, 0) + 1
            self._update_body(v)
        else:
            g = ops.CapturedGraph(lambda: self._update_body(v))
            self._graphs[v] = g
            g.replay()
        self._maybe_hard_update()
        return v

    def _update_per_epoch_prioritized(self, flush_infos):
        """Prioritised replay (no reference counterpart): per update draw rows proportionally to their
        priority, weight the loss by the importance weights and write the new |TD| priorities back.
        Round 1: eager launches (the sampler consumes host uniforms per update)."""
        rb = self.replay_buffer
        infos = []
        for _ in range(self.opt_times):
            batch = rb.random_batch(self.batch_size, self.sample_key)
            self.training_update_num += 1
            variant = self._variant()
            self._explicit_batch = batch
            try:
                self._update_body(variant)
                self._maybe_hard_update()
            finally:
                self._explicit_batch = None
            td = getattr(self, "_td", None)
            if td is not None:
                rb.update_priorities(batch["indices"], td)
            if flush_infos:
                infos.append(self._decode_info(self._ub["info"][0].cpu().numpy(), variant))
        if flush_infos:
            self._last_infos = infos
            if self.logger is not None:
                for info in infos:
                    self.logger.add_update_info(info)

    @fused.presplit_scope
    def update_per_epoch(self, flush_infos=True):
        """opt_times x {random_batch; update} (off_rl_algo.py:46-51)."""
        ub = self._ub or self._ub_setup()
        if hasattr(self.replay_buffer, "update_priorities"):
            return self._update_per_epoch_prioritized(flush_infos)
        size = self.replay_buffer.num_steps_can_sample()
        for u in range(ub["U"]):
            idx = np.random.randint(0, size, ub["b"])           # one draw per update, like random_batch
            ub["idx_host"][u * ub["b"]:(u + 1) * ub["b"]].copy_(torch.from_numpy(idx.astype(np.int64)))
        ub["idx"].copy_(ub["idx_host"], non_blocking=True)
        ub["upd"].zero_()
        variants = [self._run_update() for _ in range(ub["U"])]
        if not flush_infos:
            return
        log = ub["log32"][:len(variants)].cpu().numpy()
        self._last_infos = [self._decode_info(log[u], variants[u]) for u in range(len(variants))]
        if self.logger is not None:
            for info in self._last_infos:
                self.logger.add_update_info(info)

    @fused.presplit_scope
    def update(self, batch):
        """Eager single update on an explicit batch dict (reference signature); syncs to return floats."""
        dev = self.device
        ub = self._ub or self._ub_setup()
        conv = {}
        for k in self.sample_key:
            v = batch[k]
            v = torch.as_tensor(np.asarray(v)) if not torch.is_tensor(v) else v
            dt = torch.uint8 if k == "terminals" else torch.float32
            conv[k] = v.to(device=dev, dtype=dt).contiguous()
        self.training_update_num += 1
        variant = self._variant()
        self._explicit_batch = conv
        try:
            self._update_body(variant)
            self._maybe_hard_update()
        finally:
            self._explicit_batch = None
        ub["upd"].zero_()
        return self._decode_info(ub["info"][0].cpu().numpy(), variant)

    _explicit_batch = None

    def _batch(self):
        return self._explicit_batch if self._explicit_batch is not None else self._gather()

    def update_per_timestep(self):
        if self.replay_buffer.num_steps_can_sample() > max(self.min_pool, self.batch_size):
            self.update_per_epoch()

    def pretrain(self):
        """pretrain_epochs of collection with the learning policy, no updates (off_rl_algo.py:53-84)."""
        total_frames = 0
        self.pretrain_frames = self.pretrain_epochs * self.epoch_frames
        for pretrain_epoch in range(self.pretrain_epochs):
            start = time.time()
            self.start_epoch()
            training_epoch_info = self.collector.train_one_epoch()
            for reward in training_epoch_info["train_rewards"]:
                self.training_episode_rewards.append(reward)
            finish_epoch_info = self.finish_epoch()
            total_frames += self.epoch_frames
            infos = {"Train_Epoch_Reward": training_epoch_info["train_epoch_reward"],
                     "Running_Training_Average_Rewards":
                         np.mean(self.training_episode_rewards) if len(self.training_episode_rewards) else float("nan")}
            infos.update(finish_epoch_info)
            if self.logger is not None:
                self.logger.add_epoch_info(pretrain_epoch, total_frames, time.time() - start, infos, csv_write=False)
        if self.logger is not None:
            self.logger.log("Finished Pretrain")
