"""GAE scan microbenchmark (CUDA events, L2-exceeding working set or explicit L2 flush).

    python scripts/gae_bench.py [--sizes 128x4096,128x1048576] [--iters 20]
Prints one JSON line per (size, variant): us per launch, algorithmic GB/s (18 B/elt + 4 B*N), frac of peak.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchrl_b200 import ops  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="128x4096,128x32768,128x262144,128x1048576,1000x1024,2048x4096")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--variants", default="0,2,3,4")
    args = ap.parse_args()
    peak, how = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # 256 MB > 126 MB L2
    for sz in args.sizes.split(","):
        T, N = (int(x) for x in sz.split("x"))
        R = torch.randn(T, N, device="cuda")
        V = torch.randn(T, N, device="cuda")
        Tm = (torch.rand(T, N, device="cuda") < 0.01).to(torch.uint8)
        TL = (torch.rand(T, N, device="cuda") < 0.005).to(torch.uint8)
        LV = torch.randn(N, device="cuda")
        A = torch.empty_like(R)
        Rt = torch.empty_like(R)
        bytes_alg = 18 * T * N + 4 * N
        for variant in (int(v) for v in args.variants.split(",")):
            if (variant == 2 and N % 4) or (variant == 4 and N % 128):
                continue
            for _ in range(3):
                ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, A, Rt, variant)
            times = []
            for _ in range(args.iters):
                flush.fill_(1)  # evict L2 between timed iterations
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                ops.gae_scan(R, V, Tm, TL, LV, 0.99, 0.95, True, A, Rt, variant)
                e.record()
                torch.cuda.synchronize()
                times.append(s.elapsed_time(e) * 1e3)
            times.sort()
            med = times[len(times) // 2]
            gbs = bytes_alg / med / 1e3
            print(json.dumps({"T": T, "N": N, "variant": variant, "us_median": round(med, 2),
                              "us_min": round(times[0], 2), "GBs": round(gbs, 1),
                              "frac_of_%s_peak" % how: round(gbs / peak, 3), "l2_flush": True}), flush=True)


if __name__ == "__main__":
    main()
