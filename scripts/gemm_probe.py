"""Hardware check of the CTA-pair 3xTF32 GEMM (csrc/gemm_pair.cu) against fp64 and against the single-CTA kernel.

    python scripts/gemm_probe.py            # every case in its own subprocess under a 90 s timeout
    python scripts/gemm_probe.py case nt    # one case in this process

A wrong barrier protocol hangs the kernel: never run a case without a timeout around it.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ("nt", "nt_split", "nn", "nn_split", "tn", "time", "trace")


def timeit(fn, n=50):
    import torch
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


def planes(w):
    from torchrl_b200.networks import fused
    return fused.split_tf32(w)


SLOTS = 136


def trace_case():
    """Per-phase clock64() stamps of one CTA of the pair kernel (built with -DTRL_PAIR_TRACE into its own .so)."""
    import ctypes
    import torch
    csrc = os.path.join(ROOT, "torchrl_b200", "csrc")
    lib_path = os.path.join(ROOT, "torchrl_b200", "lib", "libtrl_pair_trace.so")
    subprocess.run(["nvcc", "-DTRL_PAIR_TRACE", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                    "-lineinfo", "-Xcompiler", "-fPIC", "-shared", os.path.join(csrc, "gemm_pair.cu"),
                    os.path.join(csrc, "errors.cu"), "-o", lib_path], check=True)
    lib = ctypes.CDLL(lib_path)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    lib.trl_gemm3_pair.argtypes = [vp, vp, vp, vp, i64, i64, i32, vp, i32, vp]
    lib.trl_gemm3_pair_tn.argtypes = [vp, vp, vp, i64, i64, i32, vp, vp]
    lib.trl_pair_set_trace.argtypes = [vp, ctypes.c_uint]
    dev = "cuda"
    M, K = 16384, 256
    a = torch.randn(M, K, device=dev)
    w = torch.randn(256, K, device=dev) / 16
    hi, lo = planes(w)
    bias = torch.randn(256, device=dev) * 0.1
    out = torch.empty(M, 256, device=dev)
    buf = torch.zeros(SLOTS, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def show(tag, fn, ctas=(0, 1, 64)):
        for cta in ctas:
            lib.trl_pair_set_trace(buf.data_ptr(), cta)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            buf.zero_()
            fn()
            torch.cuda.synchronize()
            t = buf.cpu().numpy()
            t0 = t[0]
            rel = lambda i: int(t[i] - t0) if t[i] else -1
            nkb = 8
            print("%s cta %d: setup %d  tmem_full %d  tmem->smem %d  stored %d  end %d (clocks from kernel entry)" %
                  (tag, cta, rel(1), rel(2), rel(3), rel(4), rel(5)))
            print("   tma issue   " + " ".join("%6d" % rel(8 + k) for k in range(nkb)))
            print("   full seen   " + " ".join("%6d" % rel(24 + k) for k in range(nkb)))
            print("   conv done   " + " ".join("%6d" % rel(40 + k) for k in range(nkb)))
            print("   mma start   " + " ".join("%6d" % rel(72 + k) for k in range(nkb)))
            print("   mma issued  " + " ".join("%6d" % rel(104 + k) for k in range(nkb)))
            print("   epi chunk   " + " ".join("%6d" % rel(120 + k) for k in range(8)), flush=True)

    show("nt split + tanh", lambda: lib.trl_gemm3_pair(a.data_ptr(), hi.data_ptr(), lo.data_ptr(), out.data_ptr(), M, K, 0,
                                                        bias.data_ptr(), 1, st))
    show("nt raw, no epilogue", lambda: lib.trl_gemm3_pair(a.data_ptr(), w.data_ptr(), None, out.data_ptr(), M, K, 0, None, 0,
                                                            st), ctas=(0,))
    g = torch.randn(16384, 256, device=dev)
    x = torch.randn(16384, 256, device=dev)
    ws = torch.empty(64 * 256 * 256, device=dev)
    dw = torch.empty(256, 256, device=dev)
    show("tn (wgrad) 64 splits", lambda: lib.trl_gemm3_pair_tn(g.data_ptr(), x.data_ptr(), dw.data_ptr(), 256, 16384, 64,
                                                                ws.data_ptr(), st), ctas=(0,))


def run_case(name):
    import torch
    from torchrl_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    if name in ("nt", "nt_split", "nn", "nn_split"):
        nmajor = name.startswith("nn")
        for M, K in ((256, 32), (256, 256), (1000, 256), (4096, 256), (16384, 256), (16384, 512)):
            a = torch.randn(M, K, device=dev)
            w = torch.randn(256, K, device=dev) / K ** 0.5          # forward weight (256 out, K in)
            b = w.t().contiguous() if nmajor else w                   # nn: (K, 256) row-major
            bias = torch.randn(256, device=dev) * 0.1
            ref = a.double() @ w.double().t()
            pl = planes(b) if name.endswith("split") else None
            got = ops.gemm3_pair(a, b, planes=pl, b_nmajor=nmajor)
            torch.cuda.synchronize()
            scale = ref.abs().max().item()
            err = (got.double() - ref).abs().max().item() / scale
            got_t = ops.gemm3_pair(a, b, planes=pl, b_nmajor=nmajor, bias=bias, act=1)
            err_t = (got_t.double() - torch.tanh(ref + bias.double())).abs().max().item()
            got_r = ops.gemm3_pair(a, b, planes=pl, b_nmajor=nmajor, bias=bias, act=2)
            err_r = (got_r.double() - torch.relu(ref + bias.double())).abs().max().item() / scale
            print("%-9s M=%6d K=%4d  rel err %.2e  tanh abs err %.2e  relu rel err %.2e" % (name, M, K, err, err_t, err_r),
                  flush=True)
            # 3xTF32 with fp32 TMEM accumulation: ~2e-6 of max|C| at K = 256, growing linearly with K
            tol = 5e-6 * max(1, K // 256)
            assert err < tol and err_t < tol * max(1.0, scale) and err_r < tol, "accuracy"
    elif name == "tn":
        for M, K, S in ((256, 32, 1), (256, 2048, 8), (256, 16384, 64), (512, 16384, 64)):
            g = torch.randn(K, M, device=dev)
            x = torch.randn(K, 256, device=dev)
            ref = g.double().t() @ x.double()
            got = ops.gemm3_pair_tn(g, x, splits=S)
            torch.cuda.synchronize()
            err = (got.double() - ref).abs().max().item() / ref.abs().max().item()
            print("tn        M=%6d K=%6d splits=%2d rel err %.2e" % (M, K, S, err), flush=True)
            assert err < 5e-6, "accuracy"
    elif name == "time":
        for M in (4096, 16384, 65536):
            K = 256
            a = torch.randn(M, K, device=dev)
            w = torch.randn(256, K, device=dev) / 16
            wt = w.t().contiguous()
            bias = torch.randn(256, device=dev) * 0.1
            pl, plt = planes(w), planes(wt)
            t = {}
            t["pair nt raw+tanh"] = timeit(lambda: ops.gemm3_pair(a, w, bias=bias, act=1))
            t["pair nt split+tanh"] = timeit(lambda: ops.gemm3_pair(a, w, planes=pl, bias=bias, act=1))
            t["pair nt split"] = timeit(lambda: ops.gemm3_pair(a, w, planes=pl))
            t["pair nn split"] = timeit(lambda: ops.gemm3_pair(a, w, planes=pl, b_nmajor=True))
            t["single nt+tanh"] = timeit(lambda: ops.gemm_tf32x3_nt(a, w, bias=bias, act=1))
            t["single nt"] = timeit(lambda: ops.gemm_tf32x3_nt(a, w))
            t["cublas fp32"] = timeit(lambda: torch.mm(a, wt))
            print("M=%6d K=256: " % M + "  ".join("%s %.1f us" % kv for kv in t.items()), flush=True)
        K = 16384
        g = torch.randn(K, 256, device=dev)
        x = torch.randn(K, 256, device=dev)
        ws = torch.empty(64 * 256 * 256, device=dev)
        t1 = timeit(lambda: ops.gemm3_pair_tn(g, x, splits=64, workspace=ws))
        t2 = timeit(lambda: ops.gemm_tf32x3_tn(g, x, splits=64, workspace=ws))
        t3 = timeit(lambda: torch.mm(g.t(), x))
        print("wgrad 256x256, K=16384, 64 splits (incl. reduce): pair %.1f us  single %.1f us  cublas %.1f us" % (t1, t2, t3),
              flush=True)
    elif name == "trace":
        trace_case()
    else:
        raise SystemExit("unknown case " + name)
    print("CASE %s OK" % name, flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "case":
        return run_case(sys.argv[2])
    failed = []
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "case", c], timeout=90)
            if r.returncode != 0:
                failed.append(c)
                print("CASE %s FAILED rc=%d" % (c, r.returncode), flush=True)
        except subprocess.TimeoutExpired:
            failed.append(c)
            print("CASE %s TIMED OUT (hang)" % c, flush=True)
    print("gemm_probe: failed = %s" % failed, flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
