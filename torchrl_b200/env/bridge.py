"""Device bridge for REAL host envs: a host vec env (the reference's VecEnv /
SubProcVecEnv API over NumPy, e.g. torchrl_b200.hostenv) presented with the device-env interface the
collectors drive, so the rest of the pipeline (K2 normaliser, K3 sampling, K4/K5 row store, K6 GAE,
K7-K11 update) runs unchanged on the GPU.

Stands where ``NormObs(VecEnv(...))`` stands in the reference (/root/reference/torchrl/env/get_env.py:70-87):
the per-env wrapper chain stays on the host inside each env; observation normalisation runs on the device
over the whole (N, o) batch.

One step costs exactly one D2H copy (actions, N*a floats) and one H2D copy (observations, rewards and the
two flag vectors, packed in one pinned block), both through page-locked staging memory.
"""
import numpy as np
import torch

from .. import ops
from ..spaces import is_box
from .synth import DeviceNormalizer

F32, U8, I32 = torch.float32, torch.uint8, torch.int32


class HostEnvBridge:
    host_bridge = True
    lockstep = False
    pixel = False

    def __init__(self, host_env, env_param=None, device="cuda"):
        env_param = dict(env_param or {})
        self.host = host_env
        self.env_nums = N = int(host_env.env_nums)
        self.device = dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("HostEnvBridge feeds a CUDA pipeline (there is no CPU path); got %r" % (device,))
        self.observation_space = host_env.observation_space
        self.action_space = host_env.action_space
        self.continuous = is_box(self.action_space)
        shp = tuple(self.observation_space.shape)
        if len(shp) != 1:
            raise NotImplementedError("HostEnvBridge handles flat observations; got shape %r" % (shp,))
        self.obs_dim = o = int(shp[0])
        self.act_dim = a = int(self.action_space.shape[0]) if self.continuous else 1
        self.obs_norm = bool(env_param.get("obs_norm", False))
        self._reward_scale = env_param.get("reward_scale", 1)   # informational: the host wrappers apply it
        self.training = True
        self._max_episode_steps = int(getattr(host_env, "_max_episode_steps", 0) or 0)
        # ---- staging: [obs f32 N*o | reward f32 N | done u8 N | time_limit u8 N], pinned + device twin
        nbytes = 4 * N * o + 4 * N + 2 * N
        self._pin = torch.empty(nbytes, dtype=U8, pin_memory=True)
        self._dev = torch.zeros(nbytes, dtype=U8, device=dev)
        cuts = (0, 4 * N * o, 4 * N * o + 4 * N, 4 * N * o + 5 * N, nbytes)
        pin_np = self._pin.numpy()
        self._h_obs = pin_np[cuts[0]:cuts[1]].view(np.float32).reshape(N, o)
        self._h_rew = pin_np[cuts[1]:cuts[2]].view(np.float32)
        self._h_done = pin_np[cuts[2]:cuts[3]]
        self._h_tl = pin_np[cuts[3]:cuts[4]]
        self.state = self._dev[cuts[0]:cuts[1]].view(F32).view(N, o)      # raw observation
        self.reward = self._dev[cuts[1]:cuts[2]].view(F32)
        self.done = self._dev[cuts[2]:cuts[3]]
        self.time_limit = self._dev[cuts[3]:cuts[4]]
        self._obs_cut = cuts[1]
        self.obs_out = torch.zeros(N, o, dtype=F32, device=dev)
        self._act_pin = torch.empty(N, a, dtype=F32, pin_memory=True)
        self._act_np = self._act_pin.numpy()
        self._mask_pin = torch.empty(N, dtype=U8, pin_memory=True)
        self.host_done = np.zeros(N, dtype=bool)                         # host copy of the last step's `done`
        self._obs_normalizer = DeviceNormalizer((o,), device=dev) if self.obs_norm else None
        self.dist = None
        self.h2d_bytes_per_step = nbytes
        self.d2h_bytes_per_step = 4 * N * a

    # ------------------------------------------------------------------ reference API
    def train(self):
        self.training = True
        self.host.train()

    def eval(self):
        self.training = False
        self.host.eval()

    def close(self):
        self.host.close()

    def render(self, *a, **k):
        return None

    def seed(self, seed):
        self.host.seed(seed)

    def _wait_staging(self):
        """The pinned staging block may still be the source of an H2D copy in flight (launch_step / a previous
        upload return without waiting): wait for that copy before the host writes into the block again."""
        ev = getattr(self, "_h2d_done", None)
        if ev is not None:
            ev.synchronize()

    def _mark_staging(self):
        if getattr(self, "_h2d_done", None) is None:
            self._h2d_done = torch.cuda.Event()
        self._h2d_done.record(torch.cuda.current_stream(self.device))

    def _upload_obs(self, obs):
        self._wait_staging()
        np.copyto(self._h_obs, np.asarray(obs).reshape(self._h_obs.shape), casting="same_kind")
        self._dev[:self._obs_cut].copy_(self._pin[:self._obs_cut], non_blocking=True)
        self._mark_staging()

    def _observe(self, update):
        """NormObs.observation (/root/reference/torchrl/env/base_wrapper.py:118-121) on the device."""
        if not self.obs_norm:
            self.obs_out.copy_(self.state)
            return self.obs_out
        if update and self.training:
            self._obs_normalizer.update_estimate(self.state)
        return self._obs_normalizer.filt(self.state, out=self.obs_out)

    def reset(self, **kwargs):
        self._upload_obs(self.host.reset(**kwargs))
        return self._observe(update=True)

    def partial_reset(self, index_mask, **kwargs):
        """Masked host reset; returns the RAW observations of all envs on the device (NormObs does not wrap
        partial_reset: /root/reference/torchrl/env/vecenv.py:47-51)."""
        if torch.is_tensor(index_mask):
            index_mask = index_mask.reshape(-1).bool().cpu().numpy()
        self._upload_obs(self.host.partial_reset(np.asarray(index_mask).reshape(-1).astype(bool), **kwargs))
        return self.state

    def launch_step(self, actions, *unused, **unused_kw):
        """actions (N, a) on the device -> host step -> staged results on the device; `obs_out` receives what
        env.step would return.  Synchronises the stream once (the host needs the actions)."""
        self._act_pin.copy_(actions.reshape(self._act_pin.shape), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        if self.continuous:
            acts = self._act_np
        else:
            acts = self._act_np.astype(np.int64)
        obs, rew, done, infos = self.host.step(acts)
        self._wait_staging()
        np.copyto(self._h_obs, np.asarray(obs).reshape(self._h_obs.shape), casting="same_kind")
        np.copyto(self._h_rew, np.asarray(rew).reshape(-1), casting="same_kind")
        done = np.asarray(done).reshape(-1).astype(bool)
        self._h_done[...] = done
        tl = infos.get("time_limit", None) if isinstance(infos, dict) else None
        self._h_tl[...] = 0 if tl is None else np.asarray(tl).reshape(-1).astype(bool)
        self.host_done = done
        self._dev.copy_(self._pin, non_blocking=True)
        self._mark_staging()
        return self._observe(update=True)

    def step(self, actions):
        """obs (N,o), reward (N,1), done (N,1) bool, {'time_limit': (N,) bool} -- device tensors."""
        actions = torch.as_tensor(actions, dtype=F32, device=self.device).reshape(self.env_nums, self.act_dim)
        self.launch_step(actions.contiguous())
        return self.obs_out, self.reward.unsqueeze(-1), self.done.bool().unsqueeze(-1), \
            {"time_limit": self.time_limit.bool()}

    def __deepcopy__(self, memo):
        raise TypeError("a HostEnvBridge owns live host envs and cannot be deep-copied: build a second bridge "
                        "for evaluation and pass it as eval_env")
