"""tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu) vs float64: fp32-faithful (error of the order of the cuBLAS fp32
SIMT GEMM's own error, ~1e-6 relative to max|C|), for the forward/dgrad shape and the split-K wgrad shape."""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("M,K,splits", [(128, 32, 1), (128, 256, 1), (16384, 256, 1), (300, 64, 1), (256, 16384, 64),
                                        (256, 1024, 4), (4096, 512, 2)])
def test_gemm_tf32x3_matches_fp64(M, K, splits):
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(M + K)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(256, K, device="cuda") / 8
    ref = a.double() @ b.double().t()
    out = ops.gemm_tf32x3_nt(a, b, splits=splits)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (out.double() - ref).abs().max().item() / scale
    err32 = ((a @ b.t()).double() - ref).abs().max().item() / scale
    print("M=%d K=%d splits=%d  rel err 3xTF32(tcgen05) %.2e  fp32 cuBLAS %.2e" % (M, K, splits, err, err32))
    assert err < 4 * err32 + 2e-7, (err, err32)


@pytest.mark.gpu
def test_transpose_kernel():
    import torch
    from torchrl_b200 import ops
    x = torch.randn(1000, 257, device="cuda")
    assert torch.equal(ops.transpose_f32(x), x.t().contiguous())
