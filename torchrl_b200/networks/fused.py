"""Linear + activation with the fused CUDA epilogues (csrc/mlp_epilogue.cu).

The GEMMs stay in cuBLAS (torch.mm); what is fused is everything PyTorch launches *around* them for a
hidden layer of MLPBase (/root/reference/torchrl/networks/base.py:24-44): bias add + activation in the
forward (1 launch instead of a cuBLASLt epilogue kernel + an elementwise kernel) and activation-backward +
bias-gradient reduction in the backward (1 launch instead of an elementwise kernel + a reduce kernel).
Numerically it is the same fp32 arithmetic in the same order per element (bias added to the GEMM result,
then the activation); the bias gradient is a fixed-order two-level sum.
"""
import torch
import torch.nn as nn

from .. import _lib, ops

ACT_CODES = {nn.Tanh: 1, nn.ReLU: 2}
_ENABLED = True


def set_fused_epilogue(flag):
    """Globally enable/disable the fused epilogue path (default on for CUDA inputs)."""
    global _ENABLED
    _ENABLED = bool(flag)


def fused_enabled():
    return _ENABLED


class _Workspace:
    """Per-(M,H) scratch for the backward's column-sum partials + tickets (allocated once, reused)."""
    cache = {}

    @classmethod
    def get(cls, M, H, device):
        key = (int(M), int(H), str(device))
        ws = cls.cache.get(key)
        if ws is None:
            n = int(_lib.load().trl_bias_act_bwd_scratch_floats(int(M), int(H)))
            ws = (torch.empty(max(n, 4), dtype=torch.float32, device=device),
                  torch.zeros((H + 127) // 128, dtype=torch.int32, device=device))
            cls.cache[key] = ws
        return ws


class _LinearAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, act):
        z = torch.mm(x, weight.t())
        _lib.call("trl_bias_act_fwd", z.data_ptr(), ops._chk(bias, torch.float32, "bias"), z.shape[0], z.shape[1],
                  act, ops._stream())
        ctx.save_for_backward(x, weight, z)
        ctx.act = act
        return z

    @staticmethod
    def backward(ctx, g):
        x, weight, y = ctx.saved_tensors
        g = g if g.is_contiguous() else g.contiguous()
        M, H = y.shape
        gz = torch.empty_like(y)
        db = torch.empty(H, dtype=torch.float32, device=y.device)
        scratch, tickets = _Workspace.get(M, H, y.device)
        _lib.call("trl_bias_act_bwd", g.data_ptr(), y.data_ptr(), gz.data_ptr(), db.data_ptr(), M, H, ctx.act,
                  scratch.data_ptr(), tickets.data_ptr(), ops._stream())
        dx = torch.mm(gz, weight) if ctx.needs_input_grad[0] else None
        dw = torch.mm(gz.t(), x) if ctx.needs_input_grad[1] else None
        return dx, dw, (db if ctx.needs_input_grad[2] else None), None


def linear_act(x, fc, act_code):
    """act(fc(x)) through the fused path; x may have any number of leading dims."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    out = _LinearAct.apply(x2, fc.weight, fc.bias, act_code)
    return out.reshape(tuple(lead) + (out.shape[-1],))


def can_fuse(x, fc, act_module):
    return (_ENABLED and x.is_cuda and x.dtype == torch.float32 and type(act_module) in ACT_CODES
            and fc.out_features % 4 == 0 and fc.bias is not None)
