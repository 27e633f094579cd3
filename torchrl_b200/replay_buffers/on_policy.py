"""On-policy rollout buffer: last_sample, GAE / discounted returns, minibatch iteration
(API of /root/reference/torchrl/replay_buffers/on_policy.py:5-95) on device tensors."""
import numpy as np
import torch

from .. import ops
from .base import BaseReplayBuffer


class OnPolicyReplayBufferBase:
    def last_sample(self, sample_key):
        """Row T-1 of each key (on_policy.py:9-14)."""
        return {k: getattr(self, "_" + k)[self._max_replay_buffer_size - 1] for k in sample_key}

    def _scan_inputs(self, last_value):
        lv = torch.as_tensor(last_value, dtype=torch.float32, device=self.device).reshape(-1).contiguous()
        if not hasattr(self, "_advs") or self._advs.shape != self._rewards.shape:
            self.allocate("advs", tuple(self._rewards.shape[1:]))
            self.allocate("estimate_returns", tuple(self._rewards.shape[1:]))
        return lv

    def generalized_advantage_estimation(self, last_value, gamma, tau):
        """GAE over the whole buffer (on_policy.py:16-44) -> `_advs`, `_estimate_returns`."""
        lv = self._scan_inputs(last_value)
        ops.gae_scan(self._rewards, self._values, self._terminals, self._time_limits, lv, gamma, tau,
                     self.time_limit_filter, self._advs, self._estimate_returns)

    def discount_reward(self, last_value, gamma):
        """Discounted-reward returns (on_policy.py:46-70)."""
        lv = self._scan_inputs(last_value)
        ops.discount_return(self._rewards, self._values, self._terminals, self._time_limits, lv, gamma,
                            self.time_limit_filter, self._advs, self._estimate_returns)

    def epoch_order(self, shuffle):
        """Row visiting order of one pass (on_policy.py:76-78): host NumPy, global legacy RNG."""
        if shuffle:
            return np.random.permutation(self._max_replay_buffer_size)
        return np.arange(self._max_replay_buffer_size)

    def one_iteration(self, batch_size, sample_key, shuffle):
        """Yield minibatches of batch_size//env_nums whole time-rows (on_policy.py:72-91)."""
        assert batch_size % self.env_nums == 0, "batch size should be dividable by env_nums"
        batch_size //= self.env_nums
        indices = self.epoch_order(shuffle)
        assert self._max_replay_buffer_size % batch_size == 0, \
            "rows per minibatch must divide the buffer rows (the reference's reshape fails otherwise)"
        idx = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int64)).to(self.device, non_blocking=True)
        pos = 0
        while pos < self._max_replay_buffer_size:
            yield self.gather_rows(idx[pos:pos + batch_size], sample_key)
            pos += batch_size


class OnPolicyReplayBuffer(OnPolicyReplayBufferBase, BaseReplayBuffer):
    pass
