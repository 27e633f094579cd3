"""Canonical synthetic environments -- CPU definition (TEST INFRASTRUCTURE / BASELINE ONLY).

The reference has no synthetic environment: its envs come from ``gym.make``
(/root/reference/torchrl/env/get_env.py:53) and the physics is third-party
MuJoCo.  BASELINE.json asks for "synthetic MuJoCo-shaped dynamics", so the
dynamics are *defined by this build*, once here in NumPy (float64, gym API,
one env per object -- what the reference's VecEnv / SubProcVecEnv expect) and
once in CUDA (torchrl_b200/csrc/env_step.cu, fp32, one thread per env).
PARITY UNPINNED by the reference (no counterpart); pinned GPU-vs-this-file.

Definition (``SynthMJ``; o = obs dim, a = act dim):

    params   A (o,o), B (a,o), c (o,)  drawn once from RandomState(1234+1000*o+a),
             rounded to float32 so both sides hold identical values
    step(u)  z  = s @ A + u @ B + c
             s' = RHO * s + ETA * tanh(z)
             r  = s'[0] - CTRL_COST * sum(u**2)
             elapsed += 1
             done = (|s'[1]| > term_thr)  or  elapsed >= max_episode_steps
    reset()  s[j] = INIT_SCALE * (2 * U(seed, episode, j) - 1);  episode += 1;  elapsed = 0
             U = 24-bit counter hash (murmur3 finaliser) -> exact in fp32 and fp64,
             so CPU and GPU resets agree bit-for-bit.
    obs      = s

``SynthHalfCheetah-v0``: o=17, a=6, never terminates except by the 1000-step
time limit (like HalfCheetah).  ``SynthAnt-v0``: o=111, a=8, terminates when
|s'[1]| > 2.3 (an "unhealthy" band, like Ant) or by time limit.
"""
import numpy as np

try:  # a real gym if one is installed ...
    import gym
    from gym import spaces
except ImportError:  # ... else the stand-in under oracle/shims (real package on sys.path: spawned workers re-import)
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims"))
    import gym
    from gym import spaces

RHO = 0.8
ETA = 0.5
CTRL_COST = 0.1
INIT_SCALE = 0.1
MAX_EPISODE_STEPS = 1000

SPECS = {
    # id: (obs_dim, act_dim, termination threshold on |s'[1]|)
    "SynthHalfCheetah-v0": (17, 6, float("inf")),
    "SynthAnt-v0": (111, 8, 2.3),
    # HalfCheetah-shaped with state-dependent early termination: |s'[1]| grows towards ~2.4 under this dynamics and
    # crosses 2.2 after 15-30 steps depending on the start state and the actions, so every env runs episodes of its
    # own length, ~5 % of the envs terminate on any step, and the collector's reset / bootstrap branch
    # (/root/reference/torchrl/collector/on_policy.py:132-148) has work on every step
    "SynthHalfCheetahTerm-v0": (17, 6, 2.2),
}

_M32 = 0xFFFFFFFF


def mix32(x):
    """murmur3 fmix32 on python ints / uint64 numpy arrays (values kept < 2**32)."""
    x = x & _M32
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & _M32
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & _M32
    x ^= x >> 16
    return x


def hash_uniform(seed, episode, j):
    """U(seed, episode, j) in [0,1): 24 random bits / 2**24 (exact in fp32).

    Accepts python ints or broadcastable uint64 numpy arrays.
    """
    key = (seed * 0x9E3779B1 + episode * 0x85EBCA77 + j * 0xC2B2AE3D + 0x27D4EB2F) & _M32
    return (mix32(key) >> 8) * (1.0 / 16777216.0)


def make_params(obs_dim, act_dim):
    """(A, B, c) as float32 arrays; identical on CPU and GPU."""
    rs = np.random.RandomState(1234 + 1000 * obs_dim + act_dim)
    A = (rs.randn(obs_dim, obs_dim) / np.sqrt(obs_dim)).astype(np.float32)
    B = (rs.randn(act_dim, obs_dim) / np.sqrt(act_dim)).astype(np.float32)
    c = (0.1 * rs.randn(obs_dim)).astype(np.float32)
    return A, B, c


def reset_state(seeds, episodes, obs_dim):
    """Batched reset states: seeds (n,), episodes (n,) -> (n, obs_dim) float64."""
    seeds = np.asarray(seeds, dtype=np.uint64).reshape(-1, 1)
    episodes = np.asarray(episodes, dtype=np.uint64).reshape(-1, 1)
    j = np.arange(obs_dim, dtype=np.uint64).reshape(1, -1)
    u = hash_uniform(seeds, episodes, j).astype(np.float64)
    return INIT_SCALE * (2.0 * u - 1.0)


def dynamics(s, u, A, B, c, term_thr):
    """One transition for a batch: s (n,o), u (n,a) -> s', reward (n,), done_dyn (n,)."""
    z = s @ A.astype(np.float64) + u @ B.astype(np.float64) + c.astype(np.float64)
    s2 = RHO * s + ETA * np.tanh(z)
    r = s2[:, 0] - CTRL_COST * np.sum(u * u, axis=1)
    done = np.abs(s2[:, 1]) > term_thr
    return s2, r, done


class SynthMJCore(gym.Env):
    """One synthetic MuJoCo-shaped env with the (old) gym API."""

    def __init__(self, env_id):
        self.env_id = env_id
        self.obs_dim, self.act_dim, self.term_thr = SPECS[env_id]
        self.A, self.B, self.c = make_params(self.obs_dim, self.act_dim)
        hi = np.full((self.obs_dim,), np.inf)
        self.observation_space = spaces.Box(-hi, hi)
        ub = np.ones((self.act_dim,))
        self.action_space = spaces.Box(-ub, ub)
        self._seed = 0
        self._episode = 0
        self.state = np.zeros((self.obs_dim,), dtype=np.float64)

    def seed(self, seed=None):
        self._seed = int(seed) & _M32
        self._episode = 0
        return [self._seed]

    def reset(self, **kwargs):
        self.state = reset_state([self._seed], [self._episode], self.obs_dim)[0]
        self._episode += 1
        return self.state.copy()

    def step(self, action):
        u = np.asarray(action, dtype=np.float64).reshape(1, self.act_dim)
        s2, r, done = dynamics(self.state.reshape(1, -1), u, self.A, self.B, self.c, self.term_thr)
        self.state = s2[0]
        return self.state.copy(), float(r[0]), bool(done[0]), {}


class TimeLimit(gym.Wrapper):
    """gym-0.10-style time limit.  The class *name* matters: the reference adds its
    TimeLimitAugment wrapper only if the made env's class name contains 'TimeLimit'
    (/root/reference/torchrl/env/get_env.py:54)."""

    def __init__(self, env, max_episode_steps=MAX_EPISODE_STEPS):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            done = True
        return ob, rew, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


def make_env(env_id, max_episode_steps=MAX_EPISODE_STEPS):
    return TimeLimit(SynthMJCore(env_id), max_episode_steps)


if hasattr(gym, "register") and hasattr(gym, "_REGISTRY"):
    for _eid in SPECS:
        gym.register(_eid, (lambda eid: (lambda **kw: make_env(eid, **kw)))(_eid))
