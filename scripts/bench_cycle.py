"""Where an outer FGMRES iteration spends its time: fixed-iteration solves of the scaled cylinder problem with
the coarse-level iteration count, the coarse mode and the restart length varied; the differences give the cost
of one coarse BiCGSTAB iteration, of the rest of the cycle and of the orthogonalisation."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402
from scripts.bench_solve import build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--pol", default="TM")
    ap.add_argument("--maxit", type=int, default=60)
    ap.add_argument("--sweep", type=str, default="[dict(mg_coarse_mode=1), dict(mg_coarse_mode=2)]")
    args = ap.parse_args()
    sim = build(args.N, -1e6)
    for extra in eval(args.sweep):
        sim.solver_options.update(rtol=1e-30, maxit=args.maxit)
        sim.solver_options.update(extra)
        best = None
        for rep in range(2):        # second run: preconditioner and arena cached
            try:
                getattr(sim, "solve_total_field_" + args.pol)()
            except fdfd_b200.NoConvergence as e:
                info = e.info
            best = info["seconds"] if best is None else min(best, info["seconds"])
        print(json.dumps(dict(N=args.N, pol=args.pol, opts=extra, its=info["iterations"],
                              ms_per_it=round(1e3 * best / info["iterations"], 4))), flush=True)
        for k in extra:
            sim.solver_options.pop(k, None)


if __name__ == "__main__":
    main()
