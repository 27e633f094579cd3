"""Per-env wrapper chain for a real (CPU) gym-style env, applied inline by ONE object.

The reference stacks gym wrappers: NormAct( RewardShift( TimeLimitAugment( BaseWrapper(env) ) ) )
(/root/reference/torchrl/env/get_env.py:52-67).  HostEnv folds the chain's arithmetic into one
`step` so the product does not depend on gym being importable:
  * NormAct.action        /root/reference/torchrl/env/continuous_wrapper.py:18-20
                          lb + (a + 1)/2 * (ub - lb), clipped to [lb, ub]; action_space becomes [-1, 1]
  * RewardShift.reward    /root/reference/torchrl/env/base_wrapper.py:37-41  (scaled in train mode only)
  * TimeLimitAugment.step /root/reference/torchrl/env/base_wrapper.py:152-156
                          info['time_limit'] = done and _max_episode_steps == _elapsed_steps
  * BaseWrapper.train/eval /root/reference/torchrl/env/base_wrapper.py:13-21
Observation normalisation is NOT here: with the device bridge it runs on the GPU over the whole
(N, o) batch (K2), exactly where NormObs sits in the reference (outside the vec env).
"""
import numpy as np

from ..spaces import Box, is_box


class HostEnv:
    def __init__(self, env, reward_scale=None):
        self.env = env
        self.training = True
        self._reward_scale = reward_scale
        self.observation_space = env.observation_space
        self.continuous = is_box(env.action_space)
        if self.continuous:
            self.lb = np.asarray(env.action_space.low)       # dtype as the env declares it (NormAct keeps it)
            self.ub = np.asarray(env.action_space.high)
            ub = np.ones(self.lb.shape)
            self.action_space = Box(-ub, ub)
        else:
            self.action_space = env.action_space
        # TimeLimitAugment is applied when the env carries gym's TimeLimit counters
        self._has_time_limit = hasattr(env, "_max_episode_steps")

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def seed(self, s):
        if hasattr(self.env, "seed"):
            return self.env.seed(s)

    def close(self):
        if hasattr(self.env, "close"):
            self.env.close()

    def render(self, *a, **k):
        return self.env.render(*a, **k)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        if self.continuous:
            action = np.clip(self.lb + (np.asarray(action) + 1.) * 0.5 * (self.ub - self.lb), self.lb, self.ub)
        else:
            action = int(action)
        ob, rew, done, info = self.env.step(action)
        if self._has_time_limit:
            info["time_limit"] = bool(done) and self.env._max_episode_steps == getattr(self.env, "_elapsed_steps", -1)
        if self._reward_scale is not None and self.training:
            rew = self._reward_scale * rew
        return ob, rew, done, info

    def __getattr__(self, attr):
        if attr.startswith("__") or attr == "env":
            raise AttributeError(attr)
        return getattr(self.env, attr)


def get_single_env(env_id, env_param, make=None):
    """One wrapped env (reference: get_single_env, get_env.py:52-67).  `make` defaults to gym.make;
    pass any callable `make(env_id) -> env` when gym is not installed (tests use the synthetic CPU env)."""
    if make is None:
        try:
            import gym
        except ImportError as exc:
            raise ImportError("a real host env needs gym (or pass make=...): %s" % (exc,))
        make = gym.make
    return HostEnv(make(env_id), reward_scale=(env_param or {}).get("reward_scale", None))
