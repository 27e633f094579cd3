"""Device-resident vectorised synthetic env with the reference's vec-env API.

Stands where ``NormObs(VecEnv(...))`` / ``NormObs(SubProcVecEnv(...))`` stand in the reference
(/root/reference/torchrl/env/get_env.py:70-87): same methods and attributes
(``env_nums``, ``observation_space``, ``action_space``, ``reset``, ``step``, ``partial_reset``,
``seed``, ``train``/``eval``, ``close``, ``_obs_normalizer``, ``_reward_scale``), but all N envs
advance in one CUDA launch and every returned array is a device tensor.
"""
import copy

import numpy as np
import torch

from .. import ops, _lib
from ..spaces import Box
from . import synth_spec as spec

F32, F64, U8, I32 = torch.float32, torch.float64, torch.uint8, torch.int32


class DeviceNormalizer:
    """Running mean/var observation normaliser with fp64 state on the device.

    Mirrors Normalizer (/root/reference/torchrl/env/base_wrapper.py:63-100): attributes
    ``_mean``, ``_var``, ``_count``, ``clip``, ``should_estimate``; pickles to NumPy so the
    reference's ``_obs_normalizer_{epoch}.pkl`` snapshot format keeps working
    (/root/reference/torchrl/algo/rl_algo.py:83-89).
    """

    def __init__(self, shape, clip=10., device="cuda"):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        dim = int(np.prod(self.shape))
        self._mean = torch.zeros(dim, dtype=F64, device=device)
        self._var = torch.ones(dim, dtype=F64, device=device)
        self._count = torch.full((1,), 1e-4, dtype=F64, device=device)
        self.clip = clip
        self.should_estimate = True

    def stop_update_estimate(self):
        self.should_estimate = False

    def update_estimate(self, data):
        if not self.should_estimate:
            return
        data = data.reshape(-1, self._mean.numel()).contiguous().float()
        sums = ops.obs_norm_moments(data)
        ops.obs_norm_merge(sums, data.shape[0], self._mean, self._var, self._count)

    def filt(self, raw, out=None):
        shp = raw.shape
        res = ops.obs_norm_filt(raw.reshape(-1, self._mean.numel()).contiguous().float(), self._mean, self._var,
                                self.clip, out)
        return res.reshape(shp)

    def inverse(self, raw):
        return raw * torch.sqrt(self._var).float() + self._mean.float()

    def to(self, device):
        self._mean, self._var, self._count = (t.to(device) for t in (self._mean, self._var, self._count))
        return self

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_mean"] = self._mean.cpu().numpy()
        d["_var"] = self._var.cpu().numpy()
        d["_count"] = float(self._count.item())
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        self._mean = torch.as_tensor(np.asarray(d["_mean"], dtype=np.float64)).reshape(-1).to(dev)
        self._var = torch.as_tensor(np.asarray(d["_var"], dtype=np.float64)).reshape(-1).to(dev)
        self._count = torch.full((1,), float(d["_count"]), dtype=F64, device=dev)

    def __deepcopy__(self, memo):
        new = DeviceNormalizer.__new__(DeviceNormalizer)
        new.__dict__.update({k: v for k, v in self.__dict__.items() if not torch.is_tensor(v)})
        new._mean, new._var, new._count = self._mean.clone(), self._var.clone(), self._count.clone()
        return new


class SynthVecEnv:
    """N synthetic envs on one GPU.

    env_param: {"reward_scale": float, "obs_norm": bool} (the reference's "env" config section).
    first_env / total_envs: global index range of this shard (multi-GPU: rank g owns
    [g*N, (g+1)*N) of total_envs; seeds follow VecEnv.seed with the GLOBAL index so the union
    over ranks equals the single-process env set).
    """

    def __init__(self, env_id, env_nums, env_param=None, device="cuda", first_env=0, total_envs=None,
                 max_episode_steps=spec.MAX_EPISODE_STEPS):
        env_param = dict(env_param or {})
        self.env_id = env_id
        self.env_nums = int(env_nums)
        self.device = torch.device(device)
        self.obs_dim, self.act_dim, self.term_thr = spec.SPECS[env_id]
        self.first_env = int(first_env)
        self.total_envs = int(total_envs) if total_envs is not None else self.env_nums
        self._max_episode_steps = int(max_episode_steps)
        self._reward_scale = env_param.get("reward_scale", 1)
        self.obs_norm = bool(env_param.get("obs_norm", False))
        self.training = True
        hi = np.full((self.obs_dim,), np.inf)
        self.observation_space = Box(-hi, hi)
        ub = np.ones((self.act_dim,))
        self.action_space = Box(-ub, ub)
        # True when episode boundaries are a deterministic function of the step count
        self.lockstep = not np.isfinite(self.term_thr)
        N, o, a, dev = self.env_nums, self.obs_dim, self.act_dim, self.device
        A, B, c = spec.make_params(o, a)
        self.A = torch.from_numpy(A).to(dev)
        self.B = torch.from_numpy(B).to(dev)
        self.c = torch.from_numpy(c).to(dev)
        self.lb = torch.full((a,), -1.0, dtype=F32, device=dev)
        self.ub = torch.full((a,), 1.0, dtype=F32, device=dev)
        self.state = torch.zeros(N, o, dtype=F32, device=dev)       # raw observation
        self.elapsed = torch.zeros(N, dtype=I32, device=dev)
        self.episode = torch.zeros(N, dtype=I32, device=dev)
        self.seeds = torch.zeros(N, dtype=I32, device=dev)
        self.reward = torch.zeros(N, dtype=F32, device=dev)
        self.done = torch.zeros(N, dtype=U8, device=dev)
        self.time_limit = torch.zeros(N, dtype=U8, device=dev)
        self.obs_out = torch.zeros(N, o, dtype=F32, device=dev)     # what step() returns (normalised if obs_norm)
        lib = _lib.load()
        self._nblk = int(lib.trl_synth_env_num_ctas(N))
        self._partial = torch.zeros(self._nblk, 2 * o, dtype=F64, device=dev)
        self.batch_sums = torch.zeros(2 * o, dtype=F64, device=dev)
        self._ticket = torch.zeros(1, dtype=I32, device=dev)
        self.any_reset = torch.zeros(2, dtype=I32, device=dev)
        self._obs_normalizer = DeviceNormalizer((o,), device=dev) if self.obs_norm else None
        self._obs = None
        self._host_elapsed = 0          # host mirror of `elapsed` (valid while lockstep and unperturbed)
        self._host_mirror_ok = True
        # stats merge happens in-kernel; with a DataParallelContext the batch sums are all-reduced
        # over ranks first and merged by a separate launch (global statistics)
        self._dist = None
        self._sums_red = None
        self.seed(0)

    @property
    def dist(self):
        return self._dist

    @dist.setter
    def dist(self, ctx):
        """Attach a DataParallelContext.  With PeerComm (distributed.py) the per-step batch sums move into a
        peer-mapped region so that csrc/comm.cu sums them over ranks in one small kernel (no NCCL launch per step).
        Must happen before the collector captures its step graph."""
        self._dist = ctx
        if ctx is not None and getattr(ctx, "active", False) and getattr(ctx, "peer", None) is not None:
            local, _ = ctx.peer.region("obs_sums", 8 * self.batch_sums.numel(), torch.float64)
            self.batch_sums = local[:self.batch_sums.numel()]
            self.batch_sums.zero_()
            self._sums_red = torch.zeros_like(self.batch_sums)

    def _reduce_sums(self):
        """Sum of `batch_sums` over ranks (identical on every rank)."""
        if self._sums_red is not None:
            return self._dist.peer.all_reduce_f64("obs_sums", self.batch_sums.numel(), self._sums_red)
        return self._dist.all_reduce_sum_(self.batch_sums)

    # ------------------------------------------------------------------ reference API
    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def close(self):
        pass

    def render(self, *a, **k):
        return None

    def seed(self, seed):
        _lib.call("trl_synth_env_seed", self.seeds.data_ptr(), self.episode.data_ptr(), self.env_nums,
                  int(seed) & 0xFFFFFFFF, self.total_envs & 0xFFFFFFFF, self.first_env & 0xFFFFFFFF, ops._stream())

    def _reset_kernel(self, mask):
        _lib.call("trl_synth_env_reset", self.state.data_ptr(), self.elapsed.data_ptr(), self.episode.data_ptr(),
                  self.seeds.data_ptr(), None if mask is None else ops._chk(mask, U8, "mask"), self.env_nums,
                  self.obs_dim, float(spec.INIT_SCALE), ops._stream())

    def _observe(self, update):
        """NormObs.observation (/root/reference/torchrl/env/base_wrapper.py:118-121)."""
        if not self.obs_norm:
            self.obs_out.copy_(self.state)
            return self.obs_out
        if update and self.training:
            if self.dist is not None and self.dist.active:
                ops.obs_norm_moments(self.state, self.batch_sums)
                sums = self._reduce_sums()
                nrm = self._obs_normalizer
                ops.obs_norm_merge(sums, self.total_envs, nrm._mean, nrm._var, nrm._count)
            else:
                self._obs_normalizer.update_estimate(self.state)
        return self._obs_normalizer.filt(self.state, out=self.obs_out)

    def reset(self, **kwargs):
        self._reset_kernel(None)
        self._host_elapsed = 0
        self._host_mirror_ok = True
        return self._observe(update=True)

    def partial_reset(self, index_mask, **kwargs):
        """Reset the masked envs and return the RAW observation of all envs -- the reference's
        NormObs does not wrap partial_reset, so VecEnv.partial_reset's un-normalised `_obs`
        comes back (/root/reference/torchrl/env/vecenv.py:47-51)."""
        mask = torch.as_tensor(index_mask, device=self.device).reshape(-1).to(U8).contiguous()
        self._reset_kernel(mask)
        self._host_mirror_ok = False
        return self.state

    def launch_step(self, actions, step_count=None, max_episode_frames=0, t_ptr=None):
        """Advance all envs one step: state/reward/done/time_limit staging buffers are updated and
        `obs_out` receives what env.step would return.  No host sync."""
        N, o, a = self.env_nums, self.obs_dim, self.act_dim
        update = self.obs_norm and self.training and self._obs_normalizer.should_estimate
        nrm = self._obs_normalizer
        distributed = self.dist is not None and self.dist.active
        rs = float(self._reward_scale) if self.training else 1.0
        _lib.call("trl_synth_env_step", self.state.data_ptr(), ops._chk(actions, F32, "actions"),
                  self.A.data_ptr(), self.B.data_ptr(), self.c.data_ptr(), self.lb.data_ptr(), self.ub.data_ptr(),
                  self.elapsed.data_ptr(), None if step_count is None else step_count.data_ptr(),
                  self.reward.data_ptr(), self.done.data_ptr(), self.time_limit.data_ptr(),
                  self._partial.data_ptr() if update else None, self.batch_sums.data_ptr() if update else None,
                  nrm._mean.data_ptr() if update else None, nrm._var.data_ptr() if update else None,
                  nrm._count.data_ptr() if update else None, self._ticket.data_ptr(),
                  self.any_reset.data_ptr(), None if t_ptr is None else t_ptr.data_ptr(),
                  N, o, a, spec.RHO, spec.ETA, spec.CTRL_COST,
                  float(self.term_thr) if np.isfinite(self.term_thr) else 3.0e38, rs, self._max_episode_steps,
                  int(max_episode_frames) if step_count is not None else (1 << 30),
                  1 if (update and not distributed) else 0, ops._stream())
        if update and distributed:
            ops.obs_norm_merge(self._reduce_sums(), self.total_envs, nrm._mean, nrm._var, nrm._count)
        if self.obs_norm:
            ops.obs_norm_filt(self.state, nrm._mean, nrm._var, nrm.clip, self.obs_out)
        return self.obs_out

    def step(self, actions):
        """obs (N,o), reward (N,1), done (N,1) bool, {'time_limit': (N,) bool} -- device tensors."""
        actions = torch.as_tensor(actions, dtype=F32, device=self.device).reshape(self.env_nums, self.act_dim).contiguous()
        self.launch_step(actions)
        if not self.obs_norm:
            self.obs_out.copy_(self.state)
        infos = {"time_limit": self.time_limit.bool()}
        return self.obs_out, self.reward.unsqueeze(-1), self.done.bool().unsqueeze(-1), infos

    def __deepcopy__(self, memo):
        new = SynthVecEnv.__new__(SynthVecEnv)
        for k, v in self.__dict__.items():
            if k == "_dist":
                new.__dict__[k] = v                       # the job's communicator is shared, never copied
            elif torch.is_tensor(v):
                new.__dict__[k] = v.clone()
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new
