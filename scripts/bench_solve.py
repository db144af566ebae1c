"""Time-to-solution of the scaled cylinder-scattering problem (SURVEY.md §8d inputs) on the GPU."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402


def build(N, eps, dtype="complex128", angle=45.0):
    f0 = 10e9
    lam = 299792458 / f0
    L = N / 50 * lam
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, L, L, N, N, dtype=dtype)
    sim.add_object(eps, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=40, sigma_max=10)
    sim.add_mask(80)
    sim.add_source("plane_wave", angle_deg=angle, polarization="TE")
    return sim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, nargs="+", default=[300, 1024])
    ap.add_argument("--pol", nargs="+", default=["TM"])
    ap.add_argument("--eps", type=float, default=-1e6)
    ap.add_argument("--opts", type=str, default="{}")
    ap.add_argument("--angles", type=float, nargs="+", default=[45.0], help="plane-wave angles (samples of one problem family)")
    ap.add_argument("--sweep", type=str, default="[{}]", help="list of option dicts to try")
    args = ap.parse_args()
    base = eval(args.opts)
    for N in args.N:
        for pol in args.pol:
            for extra, angle in [(e, a) for e in eval(args.sweep) for a in args.angles]:
                t0 = time.time()
                sim = build(N, args.eps, angle=angle)
                sim.solver_options.update(base)
                sim.solver_options.update(extra)
                t1 = time.time()
                try:
                    F = getattr(sim, "solve_total_field_" + pol)()
                    info = sim.solver_info[pol]
                    ok = True
                except fdfd_b200.NoConvergence as e:
                    info = e.info or {}
                    ok = False
                except Exception as e:  # noqa: BLE001
                    info = dict(error=str(e))
                    ok = False
                t2 = time.time()
                print(json.dumps(dict(N=N, pol=pol, eps=args.eps, angle=angle, opts={**base, **extra}, ok=ok, setup_s=round(t1 - t0, 2),
                                      solve_wall_s=round(t2 - t1, 3), **{k: (round(v, 4) if isinstance(v, float) and k != "relres" else v)
                                                                       for k, v in info.items()})), flush=True)


if __name__ == "__main__":
    main()
