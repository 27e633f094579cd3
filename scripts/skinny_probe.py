"""Correctness (vs float64 torch) and warm-cache timing of the skinny first / output layer kernels (csrc/skinny.cu)
at the PPO minibatch shape (M = 16384, H = 256).  Timing: each kernel alone, REPS launches inside one CUDA graph
(inputs stay in L2, as they do inside the minibatch graph), CUDA events around the replay.

    python scripts/skinny_probe.py [--M 16384] [--reps 20]
"""
import argparse
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchrl_b200 import _lib, ops  # noqa: E402
from torchrl_b200.networks import fused  # noqa: E402


NOTIME = False


def timed(fn, reps):
    if NOTIME:
        fn()
        return 0.0
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def relerr(a, b):
    return float((a.double() - b).abs().max() / (b.abs().max() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=16384)
    ap.add_argument("--H", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--notime", action="store_true", help="one launch of each kernel, no graphs (for ncu)")
    ap.add_argument("--only", type=int, default=0, help="only the first n (K, N) shapes")
    a = ap.parse_args()
    M, H = a.M, a.H
    global NOTIME
    NOTIME = a.notime
    dev = "cuda"
    torch.manual_seed(0)
    lib = _lib.load()
    st = lambda: ops._stream()
    ok = True
    shapes = ((17, 6), (17, 1), (23, 8), (5, 3), (11, 2))
    for K, N in shapes[:a.only] if a.only else shapes:
        x = torch.randn(M, K, device=dev)
        w1 = torch.randn(H, K, device=dev) / K ** 0.5
        b1 = torch.randn(H, device=dev) * 0.1
        y = torch.empty(M, H, device=dev)
        for act, name in ((1, "tanh"), (2, "relu")):
            f = lambda: _lib.call("trl_skinny_k_fwd", x.data_ptr(), w1.data_ptr(), b1.data_ptr(), y.data_ptr(), M, K, H, act, st())
            f()
            z = x.double() @ w1.double().t() + b1.double()
            ref = torch.tanh(z) if act == 1 else torch.relu(z)
            e = float((y.double() - ref).abs().max())
            t = timed(f, a.reps)
            good = e < 2e-6 * max(1.0, float(z.abs().max()))
            ok &= good
            print(f"k_fwd        K={K:2d} {name}  abs err {e:.2e}  {t:6.2f} us  {'ok' if good else 'FAIL'}")
        # first-layer backward: dW1 = (G*act'(Y))^T X, db1
        _lib.call("trl_skinny_k_fwd", x.data_ptr(), w1.data_ptr(), b1.data_ptr(), y.data_ptr(), M, K, H, 1, st())
        g = torch.randn(M, H, device=dev)
        dw = torch.empty(H, K, device=dev)
        db = torch.empty(H, device=dev)
        nscr = int(lib.trl_skinny_tn_scratch_floats(M, H, K))
        scr = torch.empty(nscr, device=dev)
        f = lambda: _lib.call("trl_skinny_act_wgrad", g.data_ptr(), y.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, H, K, 1, scr.data_ptr(), st())
        f()
        gz = g.double() * (1 - y.double() ** 2)
        e1, e2 = relerr(dw, gz.t() @ x.double()), relerr(db, gz.sum(0))
        t = timed(f, a.reps)
        good = e1 < 2e-6 and e2 < 2e-6
        ok &= good
        print(f"act_wgrad    K={K:2d}       rel err {e1:.2e} / db {e2:.2e}  {t:6.2f} us (tn + reduce)  {'ok' if good else 'FAIL'}")
        # output layer
        w3 = torch.randn(N, H, device=dev) / H ** 0.5
        b3 = torch.randn(N, device=dev) * 0.1
        o = torch.empty(M, N, device=dev)
        f = lambda: _lib.call("trl_skinny_n_fwd", y.data_ptr(), w3.data_ptr(), b3.data_ptr(), o.data_ptr(), M, H, N, st())
        f()
        e = relerr(o, y.double() @ w3.double().t() + b3.double())
        t = timed(f, a.reps)
        good = e < 2e-6
        ok &= good
        print(f"n_fwd        N={N:2d}       rel err {e:.2e}  {t:6.2f} us  {'ok' if good else 'FAIL'}")
        go = torch.randn(M, N, device=dev)
        dx = torch.empty(M, H, device=dev)
        f = lambda: _lib.call("trl_skinny_n_dgrad", go.data_ptr(), w3.data_ptr(), dx.data_ptr(), M, H, N, st())
        f()
        e = relerr(dx, go.double() @ w3.double())
        t = timed(f, a.reps)
        good = e < 2e-6
        ok &= good
        print(f"n_dgrad      N={N:2d}       rel err {e:.2e}  {t:6.2f} us  {'ok' if good else 'FAIL'}")
        db2 = torch.empty(H, device=dev)
        scr2 = torch.empty(int(lib.trl_skinny_dgrad_act_scratch_floats(M, H)), device=dev)
        f = lambda: _lib.call("trl_skinny_n_dgrad_act", go.data_ptr(), w3.data_ptr(), y.data_ptr(), dx.data_ptr(), db2.data_ptr(), M, H, N, 1, scr2.data_ptr(), st())
        f()
        gz2 = (go.double() @ w3.double()) * (1 - y.double() ** 2)
        e1, e2 = relerr(dx, gz2), relerr(db2, gz2.sum(0))
        t = timed(f, a.reps)
        good = e1 < 2e-6 and e2 < 2e-6
        ok &= good
        print(f"n_dgrad_act  N={N:2d}       rel err {e1:.2e} / db {e2:.2e}  {t:6.2f} us (+ reduce)  {'ok' if good else 'FAIL'}")
        dw3 = torch.empty(N, H, device=dev)
        db3 = torch.empty(N, device=dev)
        f = lambda: fused.skinny_tn(y, go, out=dw3, colsum=db3, out_transposed=True)
        f()
        e1, e2 = relerr(dw3, go.double().t() @ y.double()), relerr(db3, go.double().sum(0))
        t = timed(f, a.reps)
        good = e1 < 2e-6 and e2 < 2e-6
        ok &= good
        print(f"tn (dW3)     N={N:2d}       rel err {e1:.2e} / db {e2:.2e}  {t:6.2f} us (tn + reduce)  {'ok' if good else 'FAIL'}")
    # ragged M
    for Mr in (1, 37, 9000):
        x = torch.randn(Mr, 17, device=dev)
        w1 = torch.randn(H, 17, device=dev)
        b1 = torch.randn(H, device=dev)
        y = torch.empty(Mr, H, device=dev)
        _lib.call("trl_skinny_k_fwd", x.data_ptr(), w1.data_ptr(), b1.data_ptr(), y.data_ptr(), Mr, 17, H, 1, st())
        e = float((y.double() - torch.tanh(x.double() @ w1.double().t() + b1.double())).abs().max())
        w3 = torch.randn(6, H, device=dev)
        b3 = torch.randn(6, device=dev)
        o = torch.empty(Mr, 6, device=dev)
        _lib.call("trl_skinny_n_fwd", y.data_ptr(), w3.data_ptr(), b3.data_ptr(), o.data_ptr(), Mr, H, 6, st())
        e2 = relerr(o, y.double() @ w3.double().t() + b3.double())
        good = e < 1e-5 and e2 < 2e-6
        ok &= good
        print(f"ragged M={Mr}: k_fwd {e:.2e} n_fwd {e2:.2e} {'ok' if good else 'FAIL'}")
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
