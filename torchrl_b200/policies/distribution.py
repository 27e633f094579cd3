"""Tanh-squashed Gaussian on the device (API of /root/reference/torchrl/policies/distribution.py:5-79).

Sampling and log-density run in the CUDA library (csrc/collect.cu, csrc/ppo_loss.cu).  Noise
comes from one of two sources ("RNG parity"):
  * "philox"        -- counter-based Philox4x32-10 generated inside the kernel (default);
  * "reference_cpu" -- drawn on the HOST with torch's global CPU generator exactly like the
                       reference does (distribution.py:64-70) and uploaded; bit-identical
                       noise to the reference given the same torch.manual_seed.
"""
import math

import torch

from .. import ops

_NOISE_MODE = "philox"
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def set_noise_mode(mode):
    global _NOISE_MODE
    assert mode in ("philox", "reference_cpu")
    _NOISE_MODE = mode


def get_noise_mode():
    return _NOISE_MODE


def draw_reference_noise(shape, device):
    """N(0,1) from the global CPU generator, as TanhNormal.rsample does in the reference."""
    return torch.normal(torch.zeros(shape), torch.ones(shape)).to(device, non_blocking=True)


class _SampleFn(torch.autograd.Function):
    """(mean, log_std) -> (action, log_prob[, pre_tanh]) with the reparameterised gradient."""

    @staticmethod
    def forward(ctx, mean, log_std, eps, tanh_action, want_logp, rng):
        out = ops.tanh_gaussian_sample(mean, log_std, eps=eps, tanh_action=tanh_action, want_log_prob=want_logp,
                                       want_pre_tanh=True, want_eps=True, rng=rng)
        ctx.save_for_backward(out["action"], out["eps"], log_std)
        ctx.tanh_action = tanh_action
        ctx.shared_ls = (log_std.dim() == 1)
        ctx.mark_non_differentiable(out["pre_tanh"])
        lp = out["log_prob"] if want_logp else mean.new_zeros(mean.shape[:-1] + (1,))
        return out["action"], lp, out["pre_tanh"]

    @staticmethod
    def backward(ctx, g_action, g_logp, _g_pre):
        action, eps, log_std = ctx.saved_tensors
        g_mean, g_ls = ops.tanh_gaussian_sample_bwd(action, eps, log_std, g_action, g_logp, ctx.tanh_action)
        if ctx.shared_ls:
            g_ls = g_ls.reshape(-1, g_ls.shape[-1]).sum(0)
        return g_mean, g_ls, None, None, None, None


class TanhNormal:
    """X = tanh(Z), Z ~ N(mean, std).  Mirrors the reference class' methods."""

    def __init__(self, normal_mean, normal_std, epsilon=1e-6, log_std=None, rng=None):
        self.normal_mean = normal_mean
        self.normal_std = normal_std
        self.log_std = log_std if log_std is not None else torch.log(normal_std)
        self.epsilon = epsilon
        self.rng = rng

    def _eps(self):
        if _NOISE_MODE == "reference_cpu":
            return draw_reference_noise(tuple(self.normal_mean.shape), self.normal_mean.device)
        return None

    def rsample(self, return_pretanh_value=False):
        a, _, z = _SampleFn.apply(self.normal_mean, self.log_std.expand_as(self.normal_mean).contiguous()
                                  if self.log_std.dim() > 1 else self.log_std, self._eps(), True, False, self.rng)
        return (a, z) if return_pretanh_value else a

    def sample(self, return_pretanh_value=False):
        with torch.no_grad():
            return self.rsample(return_pretanh_value)

    def log_prob(self, value, pre_tanh_value=None):
        """Per-dimension log-density (distribution.py:33-45); plain torch ops (cold path)."""
        if pre_tanh_value is None:
            pre_tanh_value = torch.log((1 + value) / (1 - value)) / 2
        var = self.normal_std ** 2
        normal_lp = -((pre_tanh_value - self.normal_mean) ** 2) / (2 * var) - torch.log(self.normal_std) - _HALF_LOG_2PI
        return normal_lp - torch.log(1 - value * value + self.epsilon)

    def entropy(self):
        """Entropy of the *Normal* (not tanh-corrected), as in the reference (distribution.py:78-79)."""
        return 0.5 + _HALF_LOG_2PI + torch.log(self.normal_std)
