"""Evidence for DESIGN.md §8 (why parity of `PeriodicModeSolver2D` is asserted on converged pairs only): on its shipped example
(Periodic_Solver_2D/example_surface_wave_antenna.py: 200 x 80, TM, `eigs(A, M=B, k=8, sigma=0, v0=ones)`,
Periodic_Solver_2D.py:523-566) most of the values the UNMODIFIED reference returns are not eigenvalues of its
own pencil: their relative residuals ||A x - lambda B x|| / (||A x|| + |lambda| ||B x||) are 1e-5 ... 3e-2, and
they move when only the ARPACK start vector or the subspace size changes.  There is no reproducible answer to
be bit- or 1e-8-identical to beyond the first +- pair (fdfd_b200.PeriodicModeSolver2D returns eight converged pairs).

Runs in the build container only (needs /root/reference):   python scripts/periodic2d_reproducibility.py
"""
import json
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    mod = ref_loader.load("Periodic_Solver_2D")
    s = mod.PeriodicModeSolver2D("TM", freq=20e9, x_range=10e-3, z_range=8e-3, Nx=200, Nz=80, num_modes=8, guess=0,
                                 mode_filter=False)
    s.add_pec((2.27e-3, 2.28e-3), (1.0e-3, 2.0e-3))
    s.add_pec((0.9e-3, 1.0e-3), (0, 80))
    s.add_rectangle(10.2, 1, (1.0e-3, 2.27e-3), (0, 80), subpixels=16)
    s.add_pml(pml_width=50, n=3, sigma_max=5, direction="x+")
    # the pencil exactly as solve() builds it (Periodic_Solver_2D.py:529-554)
    m = s._effective_materials_and_masks()
    A, B = s._build_tm_system(m[0], m[2], m[4])
    free = s._free_mask(m[6], m[7], m[9], m[10])
    A, B = A[free, :][:, free].tocsc(), B[free, :][:, free].tocsc()
    n = A.shape[0]

    def run(v0, ncv=None):
        lam, X = spla.eigs(A, M=B, k=8, sigma=0, v0=v0, ncv=ncv)
        order = np.argsort(np.abs(lam))
        lam, X = lam[order], X[:, order]
        res = [float(np.linalg.norm(A @ X[:, i] - lam[i] * (B @ X[:, i])) /
                     (np.linalg.norm(A @ X[:, i]) + abs(lam[i]) * np.linalg.norm(B @ X[:, i]))) for i in range(8)]
        return lam, res

    rng = np.random.default_rng(0)
    runs = {"reference (v0 = ones)": run(np.ones(n, dtype=complex)),
            "v0 = ones, ncv = 40": run(np.ones(n, dtype=complex), ncv=40),
            "v0 = ones + 1e-3 noise": run(np.ones(n, dtype=complex) + 1e-3 * rng.standard_normal(n))}
    base = runs["reference (v0 = ones)"][0]
    for name, (lam, res) in runs.items():
        print(name)
        for i in range(8):
            move = min(abs(lam[i] - b) for b in base) / max(abs(lam[i]), 1e-300)
            print(f"   lambda = {lam[i]:.9g}   rel. residual {res[i]:.2e}   distance to the nearest reference value {move:.1e}")
    print(json.dumps({k: {"eigenvalues": [complex(x).__repr__() for x in v[0]], "residuals": v[1]} for k, v in runs.items()}))


if __name__ == "__main__":
    main()
