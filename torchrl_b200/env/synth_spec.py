"""Definition of the synthetic MuJoCo-shaped environments (constants + fixed parameters).

The reference has no synthetic env (its envs come from gym.make,
/root/reference/torchrl/env/get_env.py:53); BASELINE.json asks for synthetic
HalfCheetah-/Ant-shaped dynamics, so the build defines them:

    z  = s @ A + u @ B + c ;  s' = RHO*s + ETA*tanh(z)
    r  = s'[0] - CTRL_COST*|u|^2 ;  done = |s'[1]| > term_thr  or  elapsed >= 1000
    reset: s[j] = INIT_SCALE*(2*U(seed, episode, j) - 1),  U = 24-bit murmur3-finaliser hash

The CUDA implementation is csrc/env_step.cu; an independent NumPy float64 restatement used
as the parity oracle and as the CPU baseline's env lives in oracle/synth_env.py (tests check
that both files agree on every constant and parameter).
"""
import numpy as np

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


def make_params(obs_dim, act_dim):
    """(A, B, c) float32, drawn once from RandomState(1234 + 1000*obs_dim + act_dim)."""
    rs = np.random.RandomState(1234 + 1000 * obs_dim + act_dim)
    A = (rs.randn(obs_dim, obs_dim) / np.sqrt(obs_dim)).astype(np.float32)
    B = (rs.randn(act_dim, obs_dim) / np.sqrt(act_dim)).astype(np.float32)
    c = (0.1 * rs.randn(obs_dim)).astype(np.float32)
    return A, B, c


def is_synth(env_id):
    return env_id in SPECS
