import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchrl_b200 import ops
def bench(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
a = torch.randn(16384, 256, device="cuda"); b = torch.randn(256, 256, device="cuda")
out = torch.empty(16384, 256, device="cuda")
print("fwd/dgrad 16384x256x256: tcgen05 3xTF32 %.1f us   cuBLAS fp32 %.1f us" % (bench(lambda: ops.gemm_tf32x3_nt(a, b, out)), bench(lambda: torch.mm(a, b.t()))))
at = torch.randn(256, 16384, device="cuda"); bt = torch.randn(256, 16384, device="cuda")
o2 = torch.empty(256, 256, device="cuda")
for sp in (32, 64, 128):
    ws = torch.empty(sp * 256 * 256, device="cuda")
    print("wgrad 256x16384x256 splits=%d: tcgen05 %.1f us" % (sp, bench(lambda: ops.gemm_tf32x3_nt(at, bt, o2, splits=sp, workspace=ws))))
g = torch.randn(16384, 256, device="cuda")
print("transpose 16384x256: %.1f us" % bench(lambda: ops.transpose_f32(g)))
gz = torch.randn(16384, 256, device="cuda"); xx = torch.randn(16384, 256, device="cuda")
for sp in (32, 64, 128):
    ws = torch.empty(sp * 256 * 256, device="cuda")
    print("wgrad tn (no transposes) splits=%d: tcgen05 %.1f us" % (sp, bench(lambda: ops.gemm_tf32x3_tn(gz, xx, o2, splits=sp, workspace=ws))))
