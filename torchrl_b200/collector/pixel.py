"""Off-policy collector for the uint8 pixel env (BASELINE.json config 4): the VecCollector step
(/root/reference/torchrl/collector/base.py:184-230) with frames kept uint8 from the env to the replay ring.

Per step (one captured CUDA graph): ring write obs[t] <- frames | u8->f32 scale | epsilon-greedy Q policy |
env step (frame-stack shift + render, in place) | ring write next_obs[t] <- frames | scalar finalize
(acts / rewards / terminals / time_limits rows, episode returns, timeout + done -> reset mask) | re-render
the reset envs | ring advance.  `epsilon` lives in a device scalar refreshed from the host schedule.
"""
import numpy as np
import torch

from .. import _lib, ops
from .base import VecCollector

F32, F64, U8, I32 = torch.float32, torch.float64, torch.uint8, torch.int32


class PixelVecCollector(VecCollector):
    on_policy = False

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def _alloc_buffer(self):
        rb, N = self.replay_buffer, self._N
        rb.device = rb.device or self.device
        assert rb.env_nums == N
        frame = tuple(self.env.observation_space.shape)
        self._dedup = bool(getattr(rb, "frame_dedup", False))
        keys = [("acts", (N,), F32), ("rewards", (N, 1), F32), ("terminals", (N, 1), U8), ("time_limits", (N, 1), U8)]
        if self._dedup:
            # frame-de-duplicated ring (replay_buffers/memory_efficient.py): one frame of obs and one of next_obs per
            # row instead of two C-frame stacks
            if getattr(rb, "_stack", None) is None:
                rb.allocate_frames(frame)
        else:
            keys = [("obs", (N,) + frame, U8), ("next_obs", (N,) + frame, U8)] + keys
        for k, shp, dt in keys:
            if not hasattr(rb, "_" + k):
                rb.allocate(k, shp, dt)
        rb._ensure_device()
        self._T = rb._max_replay_buffer_size
        dev = self.device
        # the scalar finalize kernel wants (N,o)/(T,N,o) observation operands: feed it 1-wide dummies
        self._d_ob = torch.zeros(N, 1, dtype=F32, device=dev)
        self._d_state = torch.zeros(N, 1, dtype=F32, device=dev)
        self._d_rows = torch.zeros(self._T, N, 1, dtype=F32, device=dev)
        self._obs_f = torch.empty((N,) + frame, dtype=F32, device=dev)
        self._any_reset = torch.zeros(2, dtype=I32, device=dev)
        if not self._dedup:
            self._plan_obs = ops.RowCopyPlan([self.env.obs.view(1, -1)], [rb._obs], [ops.row_bytes_of(rb._obs)])
            self._plan_next = ops.RowCopyPlan([self.env.obs.view(1, -1)], [rb._next_obs], [ops.row_bytes_of(rb._next_obs)])
        self._eps_dev = torch.zeros(1, dtype=F32, device=dev)
        self._eps_host = torch.zeros(1, dtype=F32).pin_memory()

    def _o_dim(self):
        return 1

    def _step_body(self, bootstrap):
        env, rb = self.env, self.replay_buffer
        with torch.no_grad():
            if self._dedup:
                rb.write_obs(env.obs, env.elapsed)
            else:
                ops.ring_write(self._plan_obs, rb._top_dev)
            env.to_float(env.obs, self._obs_f)
            out = self.pf.explore(self._obs_f.unsqueeze(0), epsilon=self._eps_dev)
            self._act.copy_(out["action"].reshape(self._act.shape).to(F32))
            env.launch_step(self._act.reshape(-1))
            if self._dedup:
                rb.write_next_obs(env.obs)
            else:
                ops.ring_write(self._plan_next, rb._top_dev)
            _lib.call("trl_collect_finalize", self._d_ob.data_ptr(), self._d_ob.data_ptr(), self._d_state.data_ptr(),
                      self._act.data_ptr(), None, None, env.reward.data_ptr(), env.done.data_ptr(),
                      env.time_limit.data_ptr(), env.elapsed.data_ptr(), env.episode.data_ptr(), env.seeds.data_ptr(),
                      self.current_step.data_ptr(), self.train_rew.data_ptr(), self._epoch_reward.data_ptr(),
                      self._ret_log.data_ptr(), self._n_done.data_ptr(), None, None, None, self._d_ob.data_ptr(),
                      self._d_rows.data_ptr(), self._d_rows.data_ptr(), rb._acts.data_ptr(), None,
                      rb._rewards.data_ptr(), rb._terminals.data_ptr(), rb._time_limits.data_ptr(),
                      rb._top_dev.data_ptr(), self._N, 1, 1, int(self.max_episode_frames), 0.0, 0.0, 10.0, 0, 1,
                      ops._stream())
            # envs whose collector step counter was just zeroed need a fresh episode (done or timeout); the
            # finalize kernel already advanced their episode counter
            env._reset(zero_is_mask=self.current_step, episode_bias=1, bump=0)
            if hasattr(rb, "mark_inserted"):
                rb.mark_inserted()                    # prioritised ring: the new row enters with the max priority
            ops.counter_advance(None, rb._top_dev, self._T, rb._size_dev)

    def _need_bootstrap(self):
        return False

    def _step(self):
        self.pf.tick()                                   # host-side epsilon schedule -> device scalar
        self._eps_host[0] = float(self.pf.epsilon)
        self._eps_dev.copy_(self._eps_host, non_blocking=True)
        if not self.use_cuda_graph:
            self._step_body(False)
        elif False in self._graphs:
            self._graphs[False].replay()
        elif self._eager_steps < 3:
            self._eager_steps += 1
            self._step_body(False)
        else:
            g = ops.CapturedGraph(lambda: self._step_body(False))
            self._graphs[False] = g
            g.replay()
        self.replay_buffer.advance_host(1)
