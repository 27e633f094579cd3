import torch, time
torch.manual_seed(0)
M, H, K = 16384, 256, 256
def bench(f, n=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
for (H, K) in ((256, 256), (256, 17), (6, 256), (1, 256)):
    gz = torch.randn(M, H, device="cuda"); x = torch.randn(M, K, device="cuda")
    ref = gz.t().double() @ x.double()
    forms = {
        "mm(gz.t, x)": lambda: torch.mm(gz.t(), x),
        "mm(x.t, gz).t": lambda: torch.mm(x.t(), gz).t(),
        "gzT contig mm": lambda: torch.mm(gz.t().contiguous(), x),
    }
    for S in (8, 16, 32, 64, 128):
        m = M // S
        forms["bmm splitK S=%d" % S] = (lambda S=S, m=m: torch.bmm(gz.view(S, m, H).transpose(1, 2), x.view(S, m, K)).sum(0))
        forms["bmm2 splitK S=%d" % S] = (lambda S=S, m=m: torch.bmm(x.view(S, m, K).transpose(1, 2), gz.view(S, m, H)).sum(0).t())
    for name, f in forms.items():
        out = f()
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        print("H=%3d K=%3d %-22s %8.1f us  relerr %.1e" % (H, K, name, bench(f), err), flush=True)
# fwd / dgrad reference points
x = torch.randn(M, 256, device="cuda"); w = torch.randn(256, 256, device="cuda")
print("fwd mm(x, w.t) %.1f us" % bench(lambda: torch.mm(x, w.t())))
print("dgrad mm(g, w) %.1f us" % bench(lambda: torch.mm(x, w)))
