"""Prioritised replay over time rows (BASELINE.json config 4).

PARITY UNPINNED: the reference ships no prioritised replay.  This class extends
BaseReplayBuffer (/root/reference/torchrl/replay_buffers/base.py:4-54 API) with proportional
prioritisation at the buffer's own sampling granularity -- the time row: `random_batch` draws
`batch_size // env_nums` rows with probability proportional to the row priority and additionally returns
`weights` (importance weights, one per sample) and `indices` (row ids) for `update_priorities`.
Uniform numbers come from np.random.rand on the host (global legacy RNG, like the reference's
np.random.randint), so the sampled indices are reproducible against the NumPy oracle
(oracle/ref_numpy.per_sample).
"""
import numpy as np
import torch

from .. import ops
from .base import BaseReplayBuffer


class PrioritizedReplayBuffer(BaseReplayBuffer):
    def __init__(self, max_replay_buffer_size, env_nums=1, time_limit_filter=False, device=None, alpha=0.6,
                 beta=0.4, eps=1e-6):
        super().__init__(max_replay_buffer_size, env_nums, time_limit_filter, device)
        self.alpha, self.beta, self.eps = alpha, beta, eps
        self._priorities = None
        self._max_prio = None
        self._last_idx = None

    def _ensure_prio(self):
        self._ensure_device()
        if self._priorities is None:
            self._priorities = torch.zeros(self._max_replay_buffer_size, dtype=torch.float32, device=self.device)
            self._max_prio = torch.ones(1, dtype=torch.float32, device=self.device)

    def add_sample(self, sample_dict, **kwargs):
        self._ensure_device(next(iter(sample_dict.values())))
        self._ensure_prio()
        ops.per_insert(self._priorities, self._top_dev, self._max_prio)     # new rows get the max priority
        super().add_sample(sample_dict, **kwargs)

    def mark_inserted(self):
        """For collectors that write rows themselves: give the row at `_top` the running max priority."""
        self._ensure_prio()
        ops.per_insert(self._priorities, self._top_dev, self._max_prio)

    def random_batch(self, batch_size, sample_key):
        assert batch_size % self.env_nums == 0, "batch size should be dividable by env_nums"
        b = batch_size // self.env_nums
        self._ensure_prio()
        size = self.num_steps_can_sample()
        u = torch.from_numpy(np.random.rand(b)).to(self.device, non_blocking=True)
        idx, w = ops.per_sample(self._priorities, size, u, self.beta)
        out = self.gather_rows(idx, sample_key)
        out["weights"] = w.repeat_interleave(self.env_nums).unsqueeze(-1)
        out["indices"] = idx
        self._last_idx = idx
        return out

    def update_priorities(self, indices, td_errors):
        """td_errors: (b*N,) or (b*N,1) per-sample TD errors of the batch drawn with `indices`."""
        td = td_errors.reshape(indices.numel(), -1).contiguous().float()
        ops.per_update(self._priorities, indices, td, self.alpha, self.eps, self._max_prio)
