"""Minimal stand-in for the third-party `gym` package (absent in this image).

TEST INFRASTRUCTURE ONLY.  This is *not* reference code: it re-creates just the
handful of gym 0.10-era names that RchalYang/torchrl imports (Env, Wrapper,
ObservationWrapper, RewardWrapper, ActionWrapper, spaces.Box, spaces.Discrete)
so that the unmodified reference under /root/reference -- and the oracle port
in oracle/ref_port -- can be executed on CPU.  The product package
(torchrl_b200) ships its own copy of the spaces it needs and never imports this.
"""
from . import spaces  # noqa: F401


class Env:
    observation_space = None
    action_space = None
    metadata = {}

    def step(self, action):
        raise NotImplementedError

    def reset(self, **kwargs):
        raise NotImplementedError

    def render(self, mode="human"):
        return None

    def close(self):
        return None

    def seed(self, seed=None):
        return [seed]

    @property
    def unwrapped(self):
        return self


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.observation_space = getattr(env, "observation_space", None)
        self.action_space = getattr(env, "action_space", None)

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def render(self, mode="human"):
        return self.env.render(mode)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    @property
    def unwrapped(self):
        return self.env.unwrapped


class ObservationWrapper(Wrapper):
    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        return self.observation(ob), rew, done, info

    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def observation(self, observation):
        raise NotImplementedError


class RewardWrapper(Wrapper):
    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        return ob, self.reward(rew), done, info

    def reward(self, reward):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


_REGISTRY = {}


def register(env_id, factory):
    """Register a factory for `make` (used by oracle/synth_env.py)."""
    _REGISTRY[env_id] = factory


def make(env_id, **kwargs):
    if env_id not in _REGISTRY:
        # late import: the synthetic envs register themselves on import
        import importlib
        try:
            importlib.import_module("oracle.synth_env")
        except ImportError:
            pass
    if env_id not in _REGISTRY:
        raise KeyError("gym shim: unknown env id %r" % (env_id,))
    return _REGISTRY[env_id](**kwargs)
