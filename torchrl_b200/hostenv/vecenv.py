"""Host-side vectorised envs for REAL (CPU, gym-style) environments -- the callers of K1.

API of the reference's VecEnv / SubProcVecEnv
(/root/reference/torchrl/env/vecenv.py:6-78, /root/reference/torchrl/env/subproc_vecenv.py:10-157):
``VecEnv(env_nums, env_funcs, env_args)``, ``SubProcVecEnv(proc_nums, env_nums, env_funcs, env_args)``,
``reset() -> (N, o)``, ``step((N, a)) -> obs (N, o), rew (N, 1), done (N, 1) bool, {'time_limit': (N,)}``,
``partial_reset(mask (N,) bool) -> (N, o)`` (all envs' current observation, masked ones fresh),
``seed(s)`` (env i gets ``s * env_nums + i``), ``train() / eval() / close()``.

Design differences (this is the feeder of a GPU pipeline, not a port):
  * results land in PREALLOCATED per-key arrays (no per-step np.stack / dict merging); the device
    bridge (env/bridge.py) converts them into its pinned staging block with one pass per key;
  * SubProcVecEnv workers exchange actions / observations / rewards / flags through one POSIX
    shared-memory block -- the pipes carry only a command byte, nothing is pickled per step;
  * ``seed`` works in the multi-process variant (the reference's workers ignore the command,
    subproc_vecenv.py:142-144 vs :22-47).

This module imports NumPy only: spawned workers re-import it and must start quickly.
"""
import multiprocessing as mp
from multiprocessing import shared_memory

import numpy as np


def _as_lists(env_nums, env_funcs, env_args):
    if isinstance(env_funcs, (list, tuple)):
        assert len(env_funcs) == env_nums and len(env_args) == env_nums
        return list(env_funcs), list(env_args)
    return [env_funcs] * env_nums, [env_args] * env_nums


def _space_shape(space):
    shp = getattr(space, "shape", None)
    return tuple(shp) if shp is not None else ()


def _act_layout(space):
    """(width, dtype) of one env's action slot: Box -> its flat size, float64; Discrete -> 1, int64."""
    if hasattr(space, "n") and not hasattr(space, "low"):
        return 1, np.int64
    return int(np.prod(_space_shape(space))), np.float64


class _Slots:
    """The per-step exchange arrays of a vec env (views into one flat byte buffer)."""

    def __init__(self, buf, N, obs_shape, obs_dtype, act_w, act_dtype):
        self.layout = self.layout_of(N, obs_shape, obs_dtype, act_w, act_dtype)
        off = 0
        views = {}
        for name, shape, dt in self.layout:
            nb = int(np.prod(shape)) * np.dtype(dt).itemsize
            views[name] = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)), offset=off).reshape(shape)
            off += (nb + 63) & ~63
        self.obs, self.rew, self.done, self.tl, self.act, self.mask = (
            views[k] for k in ("obs", "rew", "done", "tl", "act", "mask"))

    @staticmethod
    def layout_of(N, obs_shape, obs_dtype, act_w, act_dtype):
        return [("obs", (N,) + tuple(obs_shape), obs_dtype), ("rew", (N, 1), np.float64), ("done", (N, 1), np.bool_),
                ("tl", (N,), np.bool_), ("act", (N, act_w), act_dtype), ("mask", (N,), np.bool_)]

    @classmethod
    def nbytes(cls, N, obs_shape, obs_dtype, act_w, act_dtype):
        tot = 0
        for _, shape, dt in cls.layout_of(N, obs_shape, obs_dtype, act_w, act_dtype):
            tot += (int(np.prod(shape)) * np.dtype(dt).itemsize + 63) & ~63
        return max(tot, 64)


def _step_range(envs, slots, lo, act_dtype=None):
    """Step envs[k] with slots.act[lo+k]; write results into the slots.  `act_dtype`: dtype the caller's
    action array had (the float64 slot holds float32 values exactly; the env sees the caller's dtype,
    as it does in the reference where the policy's float32 array is handed through)."""
    for k, env in enumerate(envs):
        i = lo + k
        a = slots.act[i]
        if a.dtype == np.int64:
            a = a[0] if a.shape[0] == 1 else a
        else:
            a = np.squeeze(a if act_dtype is None else a.astype(act_dtype))
        ob, r, d, info = env.step(a)
        slots.obs[i] = ob
        slots.rew[i, 0] = r
        slots.done[i, 0] = d
        slots.tl[i] = bool(info.get("time_limit", False)) if isinstance(info, dict) else False


class VecEnv:
    """N envs stepped by a loop in this process (reference: vecenv.py:6-78)."""

    def __init__(self, env_nums, env_funcs, env_args):
        self.env_nums = int(env_nums)
        self.env_funcs, self.env_args = _as_lists(self.env_nums, env_funcs, env_args)
        self.set_up_envs()

    # ------------------------------------------------------------------ construction
    def set_up_envs(self):
        self.envs = [f(*a) for f, a in zip(self.env_funcs, self.env_args)]
        self._describe(self.envs[0])
        self._buf = bytearray(_Slots.nbytes(self.env_nums, self._obs_shape, self._obs_dtype, *self._act_layout))
        self._slots = _Slots(self._buf, self.env_nums, self._obs_shape, self._obs_dtype, *self._act_layout)

    def _describe(self, example):
        self._example = example
        self._obs_shape = _space_shape(example.observation_space)
        self._obs_dtype = np.dtype(getattr(example.observation_space, "dtype", None) or np.float64)
        if self._obs_dtype.kind == "f":
            self._obs_dtype = np.dtype(np.float64)
        self._act_layout = _act_layout(example.action_space)

    @property
    def observation_space(self):
        return self._example.observation_space

    @property
    def action_space(self):
        return self._example.action_space

    @property
    def _obs(self):
        return self._slots.obs

    # ------------------------------------------------------------------ reference API
    def train(self):
        for e in self.envs:
            e.train()

    def eval(self):
        for e in self.envs:
            e.eval()

    def close(self):
        for e in self.envs:
            e.close()

    def seed(self, seed):
        for i, e in enumerate(self.envs):
            e.seed(seed * self.env_nums + i)

    def reset(self, **kwargs):
        for i, e in enumerate(self.envs):
            self._slots.obs[i] = e.reset(**kwargs)
        return self._slots.obs

    def partial_reset(self, index_mask, **kwargs):
        mask = np.asarray(index_mask).reshape(-1).astype(bool)
        for i in np.flatnonzero(mask):
            self._slots.obs[i] = self.envs[i].reset(**kwargs)
        return self._slots.obs

    def _put_actions(self, actions):
        actions = np.asarray(actions)
        self._slots.act[...] = actions.reshape(self._slots.act.shape)
        return actions.dtype.str if actions.dtype.kind == "f" and actions.dtype.itemsize < 8 else None

    def step(self, actions):
        _step_range(self.envs, self._slots, 0, self._put_actions(actions))
        return self._results()

    def _results(self):
        s = self._slots
        return s.obs, s.rew, s.done, {"time_limit": s.tl}

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        return getattr(self._example, attr)


# ======================================================================================= workers
def _worker(env_funcs, env_args, lo, N, obs_shape, obs_dtype, act_layout, shm_name, conn, parent_conn):
    parent_conn.close()
    shm = shared_memory.SharedMemory(name=shm_name)
    envs, slots = [], None
    try:
        envs = [f(*a) for f, a in zip(env_funcs, env_args)]
        slots = _Slots(shm.buf, N, obs_shape, obs_dtype, *act_layout)
        while True:
            cmd, data = conn.recv()
            if cmd == "step":
                _step_range(envs, slots, lo, data)
                conn.send(None)
            elif cmd == "reset":
                for k, e in enumerate(envs):
                    slots.obs[lo + k] = e.reset(**data)
                conn.send(None)
            elif cmd == "partial_reset":
                for k, e in enumerate(envs):
                    if slots.mask[lo + k]:
                        slots.obs[lo + k] = e.reset(**data)
                conn.send(None)
            elif cmd == "seed":
                for k, e in enumerate(envs):
                    e.seed(data + lo + k)
                conn.send(None)
            elif cmd == "train":
                for e in envs:
                    e.train()
            elif cmd == "eval":
                for e in envs:
                    e.eval()
            elif cmd == "close":
                break
    except (EOFError, KeyboardInterrupt):
        pass
    except Exception as exc:                      # surface the failure to the parent instead of hanging it
        try:
            conn.send(exc)
        except Exception:
            pass
    finally:
        for e in envs:
            try:
                e.close()
            except Exception:
                pass
        slots = None
        try:
            shm.close()
        except BufferError:
            pass
        conn.close()


class SubProcVecEnv(VecEnv):
    """N envs spread over `proc_nums` spawned worker processes (reference: subproc_vecenv.py:54-157)."""

    def __init__(self, proc_nums, env_nums, env_funcs, env_args):
        self.proc_nums = int(proc_nums)
        self._closed = False
        super().__init__(env_nums, env_funcs, env_args)

    def set_up_envs(self):
        assert self.env_nums % self.proc_nums == 0, "env_nums must be a multiple of proc_nums"
        self.env_nums_per_proc = self.env_nums // self.proc_nums
        self.example_env = self.env_funcs[0](*self.env_args[0])
        self._describe(self.example_env)
        nbytes = _Slots.nbytes(self.env_nums, self._obs_shape, self._obs_dtype, *self._act_layout)
        self._shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self._slots = _Slots(self._shm.buf, self.env_nums, self._obs_shape, self._obs_dtype, *self._act_layout)
        ctx = mp.get_context("spawn")
        self.workers, self.parent_pipes = [], []
        for p in range(self.proc_nums):
            lo, hi = p * self.env_nums_per_proc, (p + 1) * self.env_nums_per_proc
            parent, child = ctx.Pipe()
            w = ctx.Process(target=_worker, daemon=True,
                            args=(self.env_funcs[lo:hi], self.env_args[lo:hi], lo, self.env_nums, self._obs_shape,
                                  self._obs_dtype, self._act_layout, self._shm.name, child, parent))
            w.start()
            child.close()
            self.workers.append(w)
            self.parent_pipes.append(parent)

    def _roundtrip(self, cmd, data=None):
        for pipe in self.parent_pipes:
            pipe.send((cmd, data))
        for pipe in self.parent_pipes:
            res = pipe.recv()
            if isinstance(res, Exception):
                raise RuntimeError("env worker failed") from res

    def _post(self, cmd):
        for pipe in self.parent_pipes:
            pipe.send((cmd, None))

    def train(self):
        self._post("train")

    def eval(self):
        self._post("eval")

    def seed(self, seed):
        self._roundtrip("seed", int(seed) * self.env_nums)

    def reset(self, **kwargs):
        self._roundtrip("reset", kwargs)
        return self._slots.obs

    def partial_reset(self, index_mask, **kwargs):
        self._slots.mask[...] = np.asarray(index_mask).reshape(-1).astype(bool)
        self._roundtrip("partial_reset", kwargs)
        return self._slots.obs

    def step(self, actions):
        self._roundtrip("step", self._put_actions(actions))
        return self._results()

    def close(self):
        if self._closed:
            return
        self._closed = True
        for pipe in self.parent_pipes:
            try:
                pipe.send(("close", None))
            except (BrokenPipeError, OSError):
                pass
        for w in self.workers:
            w.join(timeout=5)
            if w.is_alive():
                w.terminate()
        try:
            self.example_env.close()
        except Exception:
            pass
        self._slots = None
        try:
            self._shm.close()
            self._shm.unlink()
        except (FileNotFoundError, BufferError):
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getattr__(self, attr):
        if attr.startswith("_") or attr == "example_env":
            raise AttributeError(attr)
        return getattr(self.example_env, attr)
