"""Kernel micro-benchmark: K2 apply variants, BLAS-1 and K1 at the BASELINE sizes.
Prints one line per measurement (CUDA-event timing, L2 flushed between iterations by size)."""
import argparse
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402


def timeit(fn, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, nargs="+", default=[4096])
    ap.add_argument("--dtypes", nargs="+", default=["complex128", "complex64"])
    ap.add_argument("--variants", type=int, nargs="+", default=list(range(9)))
    ap.add_argument("--pads", type=int, nargs="+", default=[0], help="extra bytes between the 5 arrays (alignment experiment)")
    args = ap.parse_args()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    for N in args.N:
        g = torch.Generator(device="cuda").manual_seed(0)
        for dtype in args.dtypes:
            td = torch.complex128 if dtype == "complex128" else torch.complex64
            rd = torch.float64 if dtype == "complex128" else torch.float32
            def rnd():
                return torch.complex(torch.rand(N, N, generator=g, device="cuda", dtype=rd) + 0.5,
                                     torch.rand(N, N, generator=g, device="cuda", dtype=rd) - 0.5)
            for pad in args.pads:
                # carve the five arrays out of one pool with `pad` extra bytes between them
                nbytes = N * N * (16 if dtype == "complex128" else 8)
                pool = torch.empty(5 * (nbytes + pad) + 4096, dtype=torch.uint8, device="cuda")
                def carve(k):
                    return pool[k * (nbytes + pad): k * (nbytes + pad) + nbytes].view(td).view(N, N)
                ca, cb, cd, x, y = (carve(k) for k in range(5))
                for t in (ca, cb, cd, x):
                    t.copy_(rnd())
                for pol in ("TE", "TM"):
                    A = fdfd_b200.HelmholtzOperator2D(pol, N, N, ca, cb, cd, 63.3, 63.3, dtype=dtype)
                    for v in args.variants:
                        A.set_variant(v)
                        med, best = timeit(lambda: A.apply(x, out=y))
                        gbs = A.bytes_per_apply / med / 1e6
                        print(json.dumps(dict(kernel="helm2d_apply", N=N, dtype=dtype, pol=pol, variant=v, pad=pad,
                                              ms=round(med, 4), best_ms=round(best, 4), GBps=round(gbs, 1),
                                              frac_of_measured_peak=round(gbs / peak, 3))), flush=True)
            x = x.clone(); y = torch.empty_like(x)
            # reference points: torch copy (the MEASURED_PEAKS methodology) and a 3-read/1-write elementwise op
            med, best = timeit(lambda: y.copy_(x))
            print(json.dumps(dict(kernel="torch_copy", N=N, dtype=dtype, ms=round(med, 4),
                                  GBps=round(2 * x.numel() * x.element_size() / med / 1e6, 1))), flush=True)
            del ca, cb, cd, x, y
        # K1 assembly
        for bc in ([0, 0], [1, 1]):
            med, best = timeit(lambda: fdfd_b200.yeeder2d_device([N, N], [0.1, 0.1], bc, [0.3, 0.7]), iters=5, warmup=2)
            print(json.dumps(dict(kernel="yee2d_assemble(DEX+DEY,+H data)", N=N, bc=bc, ms=round(med, 3))), flush=True)


if __name__ == "__main__":
    main()
