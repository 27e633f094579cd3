"""Device-resident replay ring with the reference's BaseReplayBuffer API
(/root/reference/torchrl/replay_buffers/base.py:4-54).

Storage is time-major ``(T, N, D)`` exactly like the reference (T = max_replay_buffer_size //
env_nums, each sample carries the env dimension), but lives in HBM as fp32 (uint8 for the
0/1 flags ``terminals`` / ``time_limits``) instead of float64 NumPy.  Sampling granularity is
the time ROW: indices come from the reference's own NumPy calls on the
host -- bit-exact replay indexing -- and the gather of all keys is one CUDA launch.
"""
import numpy as np
import torch

from .. import ops

FLAG_KEYS = ("terminals", "time_limits")


class BaseReplayBuffer:
    def __init__(self, max_replay_buffer_size, env_nums=1, time_limit_filter=False, device=None):
        self.env_nums = env_nums
        self._max_replay_buffer_size = int(max_replay_buffer_size) // self.env_nums
        self._top = 0
        self._size = 0
        self.time_limit_filter = time_limit_filter
        self.device = torch.device(device) if device is not None else None
        self._keys = []
        self._top_dev = None     # device mirror of _top (int32[1]) used by graph-captured writers
        self._size_dev = None
        self._gather_cache = {}

    # ------------------------------------------------------------------ storage
    def _ensure_device(self, like=None):
        if self.device is None:
            if like is not None and torch.is_tensor(like) and like.is_cuda:
                self.device = like.device
            elif torch.cuda.is_available():
                self.device = torch.device("cuda")
            else:
                raise RuntimeError("torchrl_b200 replay buffers live on the GPU (no CPU path)")
        if self._top_dev is None:
            self._top_dev = torch.full((1,), self._top, dtype=torch.int32, device=self.device)
            self._size_dev = torch.full((1,), self._size, dtype=torch.int32, device=self.device)

    def allocate(self, key, row_shape, dtype=None):
        """Create the (T,)+row_shape device tensor `_key` (zero-filled, like np.zeros in the reference)."""
        self._ensure_device()
        if dtype is None:
            dtype = torch.uint8 if key in FLAG_KEYS else torch.float32
        t = torch.zeros((self._max_replay_buffer_size,) + tuple(row_shape), dtype=dtype, device=self.device)
        setattr(self, "_" + key, t)
        if key not in self._keys:
            self._keys.append(key)
        self._gather_cache.clear()
        return t

    def _as_row(self, key, value):
        v = value
        if not torch.is_tensor(v):
            v = torch.as_tensor(np.asarray(v))
        dtype = torch.uint8 if key in FLAG_KEYS else torch.float32
        return v.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()

    def add_sample(self, sample_dict, **kwargs):
        """Write one time-row per key at `_top` and advance (base.py:19-37)."""
        self._ensure_device(next(iter(sample_dict.values())))
        rows, dsts = [], []
        for key in sample_dict:
            row = self._as_row(key, sample_dict[key])
            if not hasattr(self, "_" + key):
                self.allocate(key, tuple(row.shape))
            dst = getattr(self, "_" + key)
            if row.numel() != dst[0].numel():
                row = row.expand(dst[0].shape).contiguous()   # e.g. time_limits [False] broadcast
            rows.append(row)
            dsts.append(dst)
        for i in range(0, len(rows), 8):
            plan = ops.RowCopyPlan(rows[i:i + 8], dsts[i:i + 8], [ops.row_bytes_of(d) for d in dsts[i:i + 8]])
            ops.ring_write(plan, self._top_dev)
        ops.counter_advance(None, self._top_dev, self._max_replay_buffer_size, self._size_dev)
        self._advance()

    def terminate_episode(self):
        pass

    def _advance(self):
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def advance_host(self, steps):
        """Host mirror of `steps` device-side advances done by a captured collector graph."""
        self._top = (self._top + steps) % self._max_replay_buffer_size
        self._size = min(self._size + steps, self._max_replay_buffer_size)

    # ------------------------------------------------------------------ sampling
    def _gather_plan(self, sample_key, rows):
        ck = (tuple(sample_key), rows)
        hit = self._gather_cache.get(ck)
        if hit is not None:
            return hit
        srcs = [getattr(self, "_" + k) for k in sample_key]
        outs = [torch.empty((rows * self.env_nums,) + tuple(s.shape[2:]), dtype=s.dtype, device=self.device)
                for s in srcs]
        plans = [ops.RowCopyPlan(srcs[i:i + 8], outs[i:i + 8], [ops.row_bytes_of(s) for s in srcs[i:i + 8]])
                 for i in range(0, len(srcs), 8)]
        self._gather_cache[ck] = (plans, outs)
        return plans, outs

    def gather_rows(self, indices, sample_key, pos_ptr=None, rows=None):
        """Device gather of whole time-rows -> dict of (rows*N, D) tensors (views of cached outputs)."""
        rows = int(indices.numel()) if rows is None else int(rows)
        plans, outs = self._gather_plan(sample_key, rows)
        for p in plans:
            ops.row_gather(p, indices, rows, pos_ptr)
        return dict(zip(sample_key, outs))

    def random_batch(self, batch_size, sample_key):
        """Uniform rows with replacement (base.py:39-51): np.random.randint on the host's global
        legacy RNG (bit-exact with the reference), gather on the device."""
        assert batch_size % self.env_nums == 0, "batch size should be dividable by env_nums"
        batch_size //= self.env_nums
        size = self.num_steps_can_sample()
        indices = np.random.randint(0, size, batch_size)
        idx = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int64)).to(self.device, non_blocking=True)
        return self.gather_rows(idx, sample_key)

    def num_steps_can_sample(self):
        return self._size
