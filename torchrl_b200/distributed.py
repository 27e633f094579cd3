"""K12 -- one process per GPU, envs sharded across ranks, NCCL all-reduce of the flat gradient.

The reference is single-process (no torch.distributed / NCCL anywhere);
this is the data-parallel form of its PPO path that keeps single-process semantics:
  * rank g owns envs [g*N/G, (g+1)*N/G) -- seeds follow VecEnv.seed with the GLOBAL env index;
  * every rank draws the SAME minibatch row order (same host seed), so a global minibatch is the
    same b time-rows x all N envs as in the single-process reference;
  * the flat pf|vf gradient is summed with ONE all-reduce per minibatch and scaled by 1/G inside
    the Adam kernel, *before* global-norm clipping (clip what a single process would clip);
  * advantage-normalisation moments (ppo.py:147) and observation-normaliser batch moments
    (base_wrapper.py:75-82) are all-reduced too, so the statistics span all N envs.
Collectives are NCCL over NVLink/NVSwitch; messages are tiny (570 KB gradient, < 1 KB of moments),
i.e. latency-bound, so they are captured inside the per-minibatch / per-step CUDA graphs.
"""
import os

import torch
import torch.distributed as dist


def shard_range(total, world_size, rank):
    """[first, first+count) of a contiguous, equal-size shard; total must divide evenly."""
    if total % world_size != 0:
        raise ValueError("env count %d is not divisible by world size %d" % (total, world_size))
    per = total // world_size
    return rank * per, per


def combine_moments(sum_x, sum_x2, count):
    """(mean, unbiased std) from globally summed moments -- what torch.std() of the concatenated
    minibatch returns."""
    mean = sum_x / count
    var = (sum_x2 - sum_x * mean) / (count - 1.0)
    return mean, torch.sqrt(torch.clamp(var, min=0.0))


class DataParallelContext:
    def __init__(self, backend=None, device=None):
        self.world_size = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        self.backend = backend
        if device is None:
            device = torch.device("cuda", self.local_rank) if backend == "nccl" else torch.device("cpu")
        self.device = torch.device(device)
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        if self.world_size > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            kw = {"device_id": self.device} if backend == "nccl" else {}
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world_size, **kw)

    @property
    def active(self):
        return self.world_size > 1

    def shard(self, total_envs):
        return shard_range(total_envs, self.world_size, self.rank)

    def barrier(self):
        if self.active:
            dist.barrier()

    # ------------------------------------------------------------------ collectives
    def all_reduce_grads(self, flat_grad):
        """SUM the flat gradient over ranks; returns the scale (1/G) the optimizer must apply."""
        if not self.active:
            return 1.0
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        return 1.0 / self.world_size

    def all_reduce_sum_(self, t):
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def all_reduce_max_(self, t):
        if self.active:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def global_vec_stats(self, x, out):
        """out[0..3] = mean, unbiased std, max, min of the concatenation of `x` over all ranks (every rank holds
        the same number of elements).  CUDA: one raw-moments launch, ONE all-gather of 4 doubles per rank, one
        combine launch (capturable).  CPU tensors (gloo tests): the same arithmetic in torch ops."""
        if x.is_cuda:
            from . import _lib, ops
            if not hasattr(self, "_mom"):
                self._mom = torch.zeros(4, dtype=torch.float64, device=x.device)
                self._mom_all = torch.zeros(4 * self.world_size, dtype=torch.float64, device=x.device)
            _lib.call("trl_vec_moments", ops._chk(x, torch.float32, "x"), x.numel(), self._mom.data_ptr(), ops._stream())
            if self.active:
                dist.all_gather_into_tensor(self._mom_all, self._mom)
            else:
                self._mom_all.copy_(self._mom)
            _lib.call("trl_vec_stats_from_moments", self._mom_all.data_ptr(), self.world_size,
                      float(x.numel() * self.world_size), ops._chk(out, torch.float32, "stats"), ops._stream())
            return out
        xd = x.double()
        mom = torch.stack([xd.sum(), (xd * xd).sum()])
        ext = torch.stack([x.max(), -x.min()]).double()
        self.all_reduce_sum_(mom)
        self.all_reduce_max_(ext)
        n = float(x.numel() * self.world_size)
        mean = mom[0] / n
        var = (mom[1] - mom[0] * mean) / (n - 1.0)
        std = torch.sqrt(torch.clamp(var, min=0.0))
        out.copy_(torch.stack([mean, std, ext[0], -ext[1]]).to(out.dtype))
        return out

    def max_over_ranks(self, value):
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        self.all_reduce_max_(t)
        return float(t.item())

    def destroy(self):
        if self.active and dist.is_initialized():
            dist.destroy_process_group()
