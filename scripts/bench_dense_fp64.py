"""Library complex128 dense kernels that the 3-D solver's plane-block pivots run on: achieved
TFLOP/s of ZGEMM / inverse at the config-4 block size against the measured FP64 peak of this GPU."""
import json
import sys
import time

import torch


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 7440
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=g)
    ar, br = a.real.contiguous(), b.real.contiguous()
    t_z = timed(lambda: a @ b, 5)
    t_d = timed(lambda: ar @ br, 10)
    t_inv = timed(lambda: torch.linalg.inv(a), 3)
    # real flops: complex multiply-add = 8; LU 8/3 n^3 + inverse from LU 16/3 n^3 = 8 n^3 in total
    print(json.dumps(dict(n=n, zgemm_s=round(t_z, 5), zgemm_tflops=round(8 * n ** 3 / t_z / 1e12, 2),
                          dgemm_s=round(t_d, 5), dgemm_tflops=round(2 * n ** 3 / t_d / 1e12, 2),
                          inv_s=round(t_inv, 5), inv_tflops=round(8 * n ** 3 / t_inv / 1e12, 2))))


if __name__ == "__main__":
    main()
