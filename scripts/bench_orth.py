"""Cost of the FGMRES orthogonalisation: fixed iteration count with the (cheap) Jacobi preconditioner,
complex128 vs complex64 basis storage, at several restart lengths."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fdfd_b200.operators import HelmholtzOperator2D  # noqa: E402
from fdfd_b200.synthetic import slab_problem  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
RESTARTS = eval(sys.argv[2]) if len(sys.argv) > 2 else (10, 30, 40)
MAXIT = int(sys.argv[3]) if len(sys.argv) > 3 else 240
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ca, cb, cd, sx, sy, src, q = slab_problem(N, N, 0, N, pol="TM")
A = HelmholtzOperator2D("TM", N, N, ca, cb, cd, sx, sy)
b = torch.randn(N * N, dtype=torch.complex128, device="cuda")
for restart in RESTARTS:
    for bs in (0, 1):
        for rep in range(REPS):
            x, info = A.solve(b, method="fgmres", precond="jacobi", rtol=1e-30, maxit=MAXIT, restart=restart, basis_single=bs,
                              raise_on_noconv=False)
        print(json.dumps(dict(N=N, restart=restart, basis_single=bs, ms_per_it=round(1e3 * info["seconds"] / info["iterations"], 4),
                              iterations=info["iterations"])), flush=True)
