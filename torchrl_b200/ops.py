"""Tensor-level wrappers over the C ABI (include/torchrl_b200.h).

Each function takes torch CUDA tensors, checks dtype / contiguity / device, and launches
the sm_100a kernel on torch's *current* stream (so the calls are CUDA-graph capturable).
Memory is owned by PyTorch's caching allocator; the library never allocates.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (torchrl_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t.data_ptr()


def _tn(t):
    """(T, N) sizes of a (T,N) or (T,N,1) tensor."""
    if t.dim() == 3 and t.shape[2] == 1:
        return t.shape[0], t.shape[1]
    if t.dim() == 2:
        return t.shape[0], t.shape[1]
    raise ValueError("expected a (T,N) or (T,N,1) tensor, got %s" % (tuple(t.shape),))


def gae_scan(rewards, values, terminals, time_limits, last_value, gamma, tau, time_limit_filter,
             advs=None, returns=None, variant=1):
    """GAE backward scan (K6).  Mirrors OnPolicyReplayBufferBase.generalized_advantage_estimation
    (/root/reference/torchrl/replay_buffers/on_policy.py:16-44) on (T,N[,1]) device tensors."""
    T, N = _tn(rewards)
    if advs is None:
        advs = torch.empty_like(rewards)
    if returns is None:
        returns = torch.empty_like(rewards)
    assert values.shape == rewards.shape and terminals.shape == rewards.shape and \
        time_limits.shape == rewards.shape and last_value.numel() == N
    _lib.call("trl_gae_scan",
              _chk(rewards, torch.float32, "rewards"), _chk(values, torch.float32, "values"),
              _chk(terminals, torch.uint8, "terminals"), _chk(time_limits, torch.uint8, "time_limits"),
              _chk(last_value, torch.float32, "last_value"),
              _chk(advs, torch.float32, "advs"), _chk(returns, torch.float32, "returns"),
              T, N, float(gamma), float(tau), int(bool(time_limit_filter)), int(variant), _stream())
    return advs, returns


def discount_return(rewards, values, terminals, time_limits, last_value, gamma, time_limit_filter,
                    advs=None, returns=None, variant=1):
    """Discounted-reward returns (K6).  Mirrors OnPolicyReplayBufferBase.discount_reward
    (/root/reference/torchrl/replay_buffers/on_policy.py:46-70)."""
    T, N = _tn(rewards)
    if advs is None:
        advs = torch.empty_like(rewards)
    if returns is None:
        returns = torch.empty_like(rewards)
    assert values.shape == rewards.shape and last_value.numel() == N
    _lib.call("trl_discount_return",
              _chk(rewards, torch.float32, "rewards"), _chk(values, torch.float32, "values"),
              _chk(terminals, torch.uint8, "terminals"), _chk(time_limits, torch.uint8, "time_limits"),
              _chk(last_value, torch.float32, "last_value"),
              _chk(advs, torch.float32, "advs"), _chk(returns, torch.float32, "returns"),
              T, N, float(gamma), int(bool(time_limit_filter)), int(variant), _stream())
    return advs, returns
