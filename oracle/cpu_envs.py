"""CPU vectorised envs of the oracle port: per-env wrapper chain, in-process and multi-process
vec envs, NormObs (TEST + BASELINE INFRASTRUCTURE; see oracle/ref_port.py for the full header).

Kept free of torch imports on purpose: SubprocVecEnv workers are *spawned* (like the reference's,
/root/reference/torchrl/env/subproc_vecenv.py:7) and re-import this module, so it must be cheap.

Reference lines restated (paths under /root/reference/torchrl): env/continuous_wrapper.py:7-20 ·
env/base_wrapper.py:32-41, 103-121, 151-159 · env/get_env.py:52-87 · env/vecenv.py:6-78 ·
env/subproc_vecenv.py:10-157.
"""
import multiprocessing as mp
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import synth_env  # noqa: E402
from oracle.ref_numpy import RunningNorm  # noqa: E402


# ======================================================================================= envs
class WrappedEnv:
    """One env with the reference's per-env wrapper chain applied inline:
    NormAct( RewardShift( TimeLimitAugment( gym env ) ) )  (get_env.py:52-67)."""

    def __init__(self, env_id, env_param):
        self.core = synth_env.make_env(env_id)
        self.reward_scale = env_param.get("reward_scale", None)
        self.training = True
        self.observation_space = self.core.observation_space
        self.action_space = self.core.action_space          # already [-1, 1]
        self.lb = np.asarray(self.core.action_space.low, dtype=np.float64)
        self.ub = np.asarray(self.core.action_space.high, dtype=np.float64)

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def seed(self, s):
        self.core.seed(s)

    def close(self):
        pass

    def reset(self, **kw):
        return self.core.reset()

    def step(self, action):
        a = self.lb + (action + 1.0) * 0.5 * (self.ub - self.lb)      # NormAct.action
        a = np.clip(a, self.lb, self.ub)
        ob, rew, done, info = self.core.step(a)
        # TimeLimitAugment: flag episodes ended exactly by the time limit
        info["time_limit"] = done and self.core._max_episode_steps == self.core._elapsed_steps
        if self.reward_scale is not None and self.training:            # RewardShift (train mode only)
            rew = self.reward_scale * rew
        return ob, rew, done, info


def _make_wrapped(env_id, env_param):
    return WrappedEnv(env_id, env_param)


def _merge_infos(infos):
    keys = {}
    for d in infos:
        for k, v in d.items():
            keys.setdefault(k, []).append(v)
    return {k: np.array(v) for k, v in keys.items()}


class InProcVecEnv:
    """N envs stepped by a Python loop in this process (vecenv.py:53-61)."""

    def __init__(self, env_nums, env_id, env_param):
        self.env_nums = env_nums
        self.envs = [_make_wrapped(env_id, env_param) for _ in range(env_nums)]
        self.observation_space = self.envs[0].observation_space
        self.action_space = self.envs[0].action_space

    def train(self):
        for e in self.envs:
            e.train()

    def eval(self):
        for e in self.envs:
            e.eval()

    def close(self):
        pass

    def seed(self, seed):
        for i, e in enumerate(self.envs):
            e.seed(seed * self.env_nums + i)

    def reset(self):
        self._obs = np.stack([e.reset() for e in self.envs])
        return self._obs

    def partial_reset(self, mask):
        ids = np.argwhere(mask == 1).reshape((-1))
        self._obs[mask] = [self.envs[i].reset() for i in ids]
        return self._obs

    def step(self, actions):
        pieces = np.split(actions, self.env_nums)
        res = [e.step(np.squeeze(a)) for e, a in zip(self.envs, pieces)]
        obs, rews, dones, infos = zip(*res)
        self._obs = np.stack(obs)
        return self._obs, np.stack(rews)[:, np.newaxis], np.stack(dones)[:, np.newaxis], _merge_infos(infos)


def _worker_main(env_id, env_param, count, seeds, conn, parent_conn):
    """Command loop of one env worker (subproc_vecenv.py:10-51): its envs are stepped sequentially."""
    parent_conn.close()
    envs = [_make_wrapped(env_id, env_param) for _ in range(count)]
    try:
        while True:
            cmd, data = conn.recv()
            if cmd == "step":
                conn.send([e.step(np.squeeze(a)) for e, a in zip(envs, data)])
            elif cmd == "reset":
                conn.send([e.reset() for e in envs])
            elif cmd == "partial_reset":
                ids = np.argwhere(data == 1).reshape((-1))
                conn.send([envs[i].reset() for i in ids])
            elif cmd == "seed":
                for e, s in zip(envs, data):
                    e.seed(s)
            elif cmd == "train":
                for e in envs:
                    e.train()
            elif cmd == "eval":
                for e in envs:
                    e.eval()
            elif cmd == "close":
                conn.close()
                break
    except (EOFError, KeyboardInterrupt):
        pass


class SubprocVecEnv:
    """N envs over P spawned worker processes, one Pipe each, pickled messages, replies read in
    worker order (subproc_vecenv.py:54-157).  (The reference's `seed` command is never handled by
    its workers -- SURVEY.md A.3 -- so its subprocess envs always run unseeded; this port does
    handle it so that the baseline is reproducible; timing is unaffected.)"""

    def __init__(self, proc_nums, env_nums, env_id, env_param):
        assert env_nums % proc_nums == 0
        self.env_nums, self.proc_nums = env_nums, proc_nums
        self.per = env_nums // proc_nums
        example = _make_wrapped(env_id, env_param)
        self.observation_space, self.action_space = example.observation_space, example.action_space
        ctx = mp.get_context("spawn")
        self.pipes, self.procs = [], []
        for _ in range(proc_nums):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker_main, args=(env_id, env_param, self.per, None, child, parent), daemon=True)
            p.start()
            child.close()
            self.pipes.append(parent)
            self.procs.append(p)

    def _broadcast(self, cmd, payloads=None):
        for i, pipe in enumerate(self.pipes):
            pipe.send((cmd, None if payloads is None else payloads[i]))

    def train(self):
        self._broadcast("train")

    def eval(self):
        self._broadcast("eval")

    def close(self):
        self._broadcast("close")
        for p in self.procs:
            p.join(timeout=5)

    def seed(self, seed):
        self._broadcast("seed", [[seed * self.env_nums + w * self.per + j for j in range(self.per)]
                                 for w in range(self.proc_nums)])

    def reset(self):
        self._broadcast("reset")
        obs = []
        for pipe in self.pipes:
            obs += pipe.recv()
        self._obs = np.stack(obs)
        return self._obs

    def partial_reset(self, mask):
        self._broadcast("partial_reset", np.split(mask, self.proc_nums))
        got = []
        for pipe in self.pipes:
            got += pipe.recv()
        self._obs[mask] = got
        return self._obs

    def step(self, actions):
        pieces = np.split(actions, self.env_nums)
        self._broadcast("step", [pieces[w * self.per:(w + 1) * self.per] for w in range(self.proc_nums)])
        res = []
        for pipe in self.pipes:
            res += pipe.recv()
        obs, rews, dones, infos = zip(*res)
        self._obs = np.stack(obs)
        return self._obs, np.stack(rews)[:, np.newaxis], np.stack(dones)[:, np.newaxis], _merge_infos(infos)


class NormObsVec:
    """NormObs over a vec env (base_wrapper.py:103-121): update stats then filter, training mode only;
    partial_reset is NOT wrapped, so raw observations come back after a partial reset (quirk A.1)."""

    def __init__(self, venv):
        self.venv = venv
        self.env_nums = venv.env_nums
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.norm = RunningNorm(venv.observation_space.shape[0])
        self.training = True

    def _filter(self, ob):
        if self.training:
            self.norm.update(ob)
        return self.norm.filt(ob)

    def train(self):
        self.training = True
        self.venv.train()

    def eval(self):
        self.training = False
        self.venv.eval()

    def close(self):
        self.venv.close()

    def seed(self, s):
        self.venv.seed(s)

    def reset(self):
        return self._filter(self.venv.reset())

    def partial_reset(self, mask):
        return self.venv.partial_reset(mask)

    def step(self, actions):
        ob, r, d, info = self.venv.step(actions)
        return self._filter(ob), r, d, info


def make_vec_env(env_id, env_param, env_nums, proc_nums=0):
    """get_vec_env / get_subprocvec_env (get_env.py:70-87)."""
    venv = SubprocVecEnv(proc_nums, env_nums, env_id, env_param) if proc_nums else \
        InProcVecEnv(env_nums, env_id, env_param)
    if env_param.get("obs_norm"):
        venv = NormObsVec(venv)
    return venv


