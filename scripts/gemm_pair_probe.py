"""Round-2 starting point: build and check the WORK-IN-PROGRESS CTA-pair 3xTF32 GEMM (csrc/wip/gemm_tf32x3_pair.cu,
never run on hardware yet) against fp64 and against the production single-CTA kernel.

    timeout 120 python scripts/gemm_pair_probe.py          # ALWAYS under a short timeout: a wrong barrier protocol hangs

Steps: nvcc -> torchrl_b200/lib/libtrl_wip.so; tiny shape first (M=256, K=32), then the MLP shape (M=16384, K=256).
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from torchrl_b200 import ops  # noqa: E402

SRC = os.path.join(ROOT, "torchrl_b200", "csrc", "wip", "gemm_tf32x3_pair.cu")
LIB = os.path.join(ROOT, "torchrl_b200", "lib", "libtrl_wip.so")


def build():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "-shared", "-cudart", "static", SRC, "-o", LIB, "-I" + os.path.join(ROOT, "include")]
    subprocess.run(cmd, check=True)
    lib = ctypes.CDLL(LIB)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.trl_wip_gemm_tf32x3_nt_pair.argtypes = [vp, vp, vp, i64, i64, vp, i32, vp]
    lib.trl_wip_gemm_tf32x3_nt_pair.restype = i32
    return lib


def pair(lib, a, b, bias=None, act=0):
    out = torch.empty(a.shape[0], 256, device=a.device)
    rc = lib.trl_wip_gemm_tf32x3_nt_pair(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], a.shape[1],
                                         None if bias is None else bias.data_ptr(), act,
                                         torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    return out


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


def main():
    lib = build()
    torch.manual_seed(0)
    for M, K in ((256, 32), (256, 256), (1024, 256), (16384, 256)):
        a = torch.randn(M, K, device="cuda")
        b = torch.randn(256, K, device="cuda") / K ** 0.5
        bias = torch.randn(256, device="cuda") * 0.1
        ref = a.double() @ b.double().t()
        got = pair(lib, a, b)
        torch.cuda.synchronize()
        scale = ref.abs().max().item()
        err = (got.double() - ref).abs().max().item() / scale
        err1 = (ops.gemm_tf32x3_nt(a, b).double() - ref).abs().max().item() / scale
        got_t = pair(lib, a, b, bias, 1)
        err_t = (got_t.double() - torch.tanh(ref + bias.double())).abs().max().item()
        print("M=%6d K=%4d  rel err pair %.2e  single %.2e  tanh epilogue abs err %.2e" % (M, K, err, err1, err_t), flush=True)
        assert err < 5e-6 and err_t < 5e-6
    t_pair = timeit(lambda: pair(lib, a, b, bias, 1))
    t_one = timeit(lambda: ops.gemm_tf32x3_nt(a, b, bias=bias, act=1))
    print("M=16384 K=256 fused bias+tanh: pair %.1f us   single-CTA %.1f us (warm L2, back-to-back launches)" % (t_pair, t_one))


if __name__ == "__main__":
    main()
