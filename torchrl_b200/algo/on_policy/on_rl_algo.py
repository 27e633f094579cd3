"""On-policy base: GAE over the epoch's rollout, then minibatch updates
(API of /root/reference/torchrl/algo/on_policy/on_rl_algo.py:6-48)."""
import torch

from ..rl_algo import RLAlgo


class OnRLAlgo(RLAlgo):
    def __init__(self, shuffle=True, tau=None, gae=True, **kwargs):
        super().__init__(**kwargs)
        self.sample_key = ["obs", "acts", "advs", "estimate_returns"]
        self.shuffle = shuffle
        self.tau = tau
        self.gae = gae

    def process_epoch_samples(self):
        """last_value = V(next_obs[T-1]) * (1 - terminals[T-1]); then GAE or discounted returns
        (on_rl_algo.py:22-33).  All on the device, no host copy."""
        sample = self.replay_buffer.last_sample(['next_obs', 'terminals', 'time_limits'])
        with torch.no_grad():
            last_value = self.vf(sample['next_obs']).reshape(-1)
            last_value = (last_value * (1.0 - sample['terminals'].reshape(-1).float())).contiguous()
        if self.gae:
            self.replay_buffer.generalized_advantage_estimation(last_value, self.discount, self.tau)
        else:
            self.replay_buffer.discount_reward(last_value, self.discount)

    def update_per_epoch(self):
        self.process_epoch_samples()
        for batch in self.replay_buffer.one_iteration(self.batch_size, self.sample_key, self.shuffle):
            infos = self.update(batch)
            self.logger.add_update_info(infos)

    @property
    def networks(self):
        return [self.pf, self.vf]
