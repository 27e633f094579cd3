"""Fused bias+activation epilogue (csrc/mlp_epilogue.cu) vs the plain PyTorch layer sequence of MLPBase
(/root/reference/torchrl/networks/base.py:24-44).  Same fp32 formula per element; tolerance covers the
tanhf-vs-torch.tanh implementation difference (1e-6) and the bias-gradient summation order (1e-5 rel)."""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("act", ["tanh", "relu"])
@pytest.mark.parametrize("M,inp,hidden", [(16384, 17, (256, 256)), (300, 23, (64, 400, 300)), (5, 17, (8,))])
def test_fused_mlp_matches_plain_torch(act, M, inp, hidden):
    import copy
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    from torchrl_b200.networks import fused
    torch.manual_seed(0)
    A = nn.Tanh if act == "tanh" else nn.ReLU
    net = networks.Net(input_shape=inp, output_shape=6, hidden_shapes=list(hidden), append_hidden_shapes=[],
                       base_type=networks.MLPBase, activation_func=A).cuda()
    ref = copy.deepcopy(net)
    x = torch.randn(M, inp, device="cuda")
    w = torch.randn(M, 6, device="cuda")
    fused.set_fused_epilogue(True)
    y1 = net(x)
    (y1 * w).sum().backward()
    fused.set_fused_epilogue(False)
    try:
        y0 = ref(x)
        (y0 * w).sum().backward()
    finally:
        fused.set_fused_epilogue(True)
    torch.testing.assert_close(y1, y0, rtol=1e-5, atol=1e-6)
    for (n1, p1), (n0, p0) in zip(net.named_parameters(), ref.named_parameters()):
        scale = p0.grad.abs().max().item() + 1e-12
        torch.testing.assert_close(p1.grad, p0.grad, rtol=1e-4, atol=1e-5 * scale, msg=n1)


@pytest.mark.gpu
def test_fused_path_handles_leading_dims_and_no_grad():
    import torch
    import torch.nn as nn
    import torchrl_b200.networks as networks
    net = networks.MLPBase(input_shape=17, hidden_shapes=[32, 32], activation_func=nn.Tanh).cuda()
    x = torch.randn(1, 40, 17, device="cuda")
    with torch.no_grad():
        y = net(x)
    assert y.shape == (1, 40, 32)
    torch.testing.assert_close(y, net.seq_fcs(x), rtol=1e-5, atol=1e-6)
