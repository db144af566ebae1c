"""Micro-benchmark of the K4 multi-vector kernels (Gram-Schmidt building blocks)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from fdfd_b200.eigs import DeviceBlas  # noqa: E402

n = 4096 * 4096
blas = DeviceBlas("cuda")
PADS = eval(sys.argv[1]) if len(sys.argv) > 1 else (0,)
for dt, eb, pad in [(d, e, p_) for (d, e) in ((torch.complex128, 16), (torch.complex64, 8)) for p_ in PADS]:
    V = torch.randn((9, n + pad), dtype=dt, device="cuda")[:, :n]
    w = torch.randn(n, dtype=dt, device="cuda")
    for k in (1, 4, 8) if len(PADS) == 1 else (8,):
        for name in ("multi_dot", "multi_axpy"):
            h = np.ones(k, dtype=complex) * 1e-3
            fn = (lambda: blas.multi_dot(V, k, w)) if name == "multi_dot" else (lambda: blas.multi_axpy(V, k, w, h, -1.0, want_norm=True))
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            reps = 20
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            dt_s = (time.perf_counter() - t) / reps
            bytes_ = n * (k * eb + (eb if name == "multi_dot" else 2 * eb))
            print(json.dumps(dict(kernel=name, dtype=str(dt), pad=pad, k=k, us=round(dt_s * 1e6, 1), GBps=round(bytes_ / dt_s / 1e9, 1))), flush=True)
