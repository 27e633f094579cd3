"""Epoch loop shared by every agent (API of /root/reference/torchrl/algo/rl_algo.py:14-190)."""
import os.path as osp
import pathlib
import pickle
import time
from collections import deque

import numpy as np
import torch

from ..spaces import is_box
from .. import ops
from . import utils as atu


class _LRGroup(dict):
    """param_group stand-in: assigning ['lr'] forwards to the flat optimizer's device scalar."""

    def __init__(self, owner, seg, lr):
        super().__init__(lr=lr)
        self._owner, self._seg = owner, seg

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        if k == "lr":
            self._owner.set_lr(self._seg, v)


class SegmentOptimizer:
    """What `agent.pf_optimizer` etc. return: one segment of the agent's FlatAdam, exposing the
    `param_groups[...]['lr']` knob that update_linear_schedule uses."""

    def __init__(self, flat, seg):
        self.flat, self.seg = flat, seg
        self.param_groups = [_LRGroup(flat, seg, flat.initial_lrs[seg])]

    def zero_grad(self):
        self.flat.grad[self.flat.seg_begin[self.seg]:self.flat.seg_begin[self.seg + 1]].zero_()

    def step(self):
        self.flat.step(active_mask=1 << self.seg)


class RLAlgo:
    def __init__(self, env=None, replay_buffer=None, collector=None, logger=None, grad_clip=None, discount=0.99,
                 num_epochs=3000, batch_size=128, device='cpu', save_interval=100, eval_interval=1, save_dir=None,
                 use_cuda_graph=True, dist=None, resume_checkpoints=False):
        self.env = env
        self.continuous = is_box(self.env.action_space)
        self.replay_buffer = replay_buffer
        self.collector = collector
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("torchrl_b200 agents run on a CUDA device (there is no CPU path); got %r" % (device,))
        self.discount = discount
        self.num_epochs = num_epochs
        self.epoch_frames = self.collector.epoch_frames
        self.batch_size = batch_size
        self.training_update_num = 0
        self.sample_key = None
        self.grad_clip = grad_clip
        self.logger = logger
        self.episode_rewards = deque(maxlen=30)
        self.training_episode_rewards = deque(maxlen=30)
        self.save_interval = save_interval
        self.save_dir = save_dir
        if self.save_dir is not None:
            pathlib.Path(self.save_dir).mkdir(parents=True, exist_ok=True)
        self.best_eval = None
        self.eval_interval = eval_interval
        self.explore_time = 0
        self.train_time = 0
        self.start = time.time()
        self.use_cuda_graph = bool(use_cuda_graph)
        self.dist = dist                    # None or a torchrl_b200.distributed.DataParallelContext
        self.current_epoch = 0
        self.first_epoch = 0                # load_checkpoint moves it past the last finished epoch
        self.resume_checkpoints = bool(resume_checkpoints)

    # ------------------------------------------------------------------ resume (absent in the reference)
    def save_checkpoint(self, path):
        """Full training state (weights, Adam moments, targets, collector / env / normaliser state, replay
        ring, every RNG stream): see utils/checkpoint.py."""
        from ..utils.checkpoint import save_checkpoint
        return save_checkpoint(self, path)

    def load_checkpoint(self, path):
        """Restore a `save_checkpoint` file in place; `train()` then continues with the next epoch."""
        from ..utils.checkpoint import load_checkpoint
        self.first_epoch = load_checkpoint(self, path) + 1
        return self.first_epoch

    def start_epoch(self):
        pass

    def finish_epoch(self):
        return {}

    def pretrain(self):
        pass

    def update_per_epoch(self):
        pass

    def snapshot(self, prefix, epoch):
        """model_{name}_{epoch}.pth state_dicts + pickled obs normaliser (rl_algo.py:83-94)."""
        if prefix is None:
            return
        if hasattr(self.env, "_obs_normalizer") and self.env._obs_normalizer is not None:
            with open(osp.join(prefix, "_obs_normalizer_{}.pkl".format(epoch)), "wb") as f:
                pickle.dump(self.env._obs_normalizer, f)
        for name, network in self.snapshot_networks:
            torch.save(network.state_dict(), osp.join(prefix, "model_{}_{}.pth".format(name, epoch)))

    def train(self):
        self.pretrain()
        total_frames = 0
        if hasattr(self, "pretrain_frames"):
            total_frames = self.pretrain_frames
        self.start_epoch()
        for epoch in range(self.first_epoch, self.num_epochs):
            self.current_epoch = epoch
            self.start_epoch()

            t0 = time.time()
            training_epoch_info = self.collector.train_one_epoch()
            for reward in training_epoch_info["train_rewards"]:
                self.training_episode_rewards.append(reward)
            self.explore_time += time.time() - t0

            t0 = time.time()
            self.update_per_epoch()
            torch.cuda.synchronize(self.device)
            self.train_time += time.time() - t0

            finish_epoch_info = self.finish_epoch()
            total_frames += self.epoch_frames

            if epoch % self.eval_interval == 0:
                t0 = time.time()
                eval_infos = self.collector.eval_one_epoch()
                eval_time = time.time() - t0
                infos = {}
                for reward in eval_infos["eval_rewards"]:
                    self.episode_rewards.append(reward)
                mean_eval = np.mean(eval_infos["eval_rewards"])
                if self.best_eval is None or mean_eval > self.best_eval:
                    self.best_eval = mean_eval
                    self.snapshot(self.save_dir, 'best')
                del eval_infos["eval_rewards"]
                infos["Running_Average_Rewards"] = np.mean(self.episode_rewards)
                infos["Train_Epoch_Reward"] = training_epoch_info["train_epoch_reward"]
                infos["Running_Training_Average_Rewards"] = \
                    np.mean(self.training_episode_rewards) if len(self.training_episode_rewards) else float("nan")
                infos["Explore_Time"] = self.explore_time
                infos["Train___Time"] = self.train_time
                infos["Eval____Time"] = eval_time
                self.explore_time = 0
                self.train_time = 0
                infos.update(eval_infos)
                infos.update(finish_epoch_info)
                self.logger.add_epoch_info(epoch, total_frames, time.time() - self.start, infos)
                self.start = time.time()

            if epoch % self.save_interval == 0:
                self.snapshot(self.save_dir, epoch)
                if self.resume_checkpoints and self.save_dir is not None:
                    self.save_checkpoint(osp.join(self.save_dir, "checkpoint_latest.pt"))

        self.snapshot(self.save_dir, "finish")
        self.collector.terminate()
        self.logger.finish()

    def update(self, batch):
        raise NotImplementedError

    def _update_target_networks(self):
        """Polyak update of the target nets (rl_algo.py:169-172) on the flat buffers: one launch,
        captured inside the update graph.  The periodic HARD copy (rl_algo.py:173-176) depends on a
        host counter, so it is applied by `_maybe_hard_update` outside of any captured graph."""
        if self.use_soft_update:
            ops.polyak_update(self._target_flat.data, self._target_source(), self.tau)

    def _maybe_hard_update(self):
        if not self.use_soft_update and self.training_update_num % self.target_hard_update_period == 0:
            self._target_flat.data.copy_(self._target_source())

    @property
    def networks(self):
        return []

    @property
    def snapshot_networks(self):
        return []

    @property
    def target_networks(self):
        return []

    def to(self, device):
        for net in self.networks:
            net.to(device)
