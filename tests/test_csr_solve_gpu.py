"""Persistent BiCGSTAB kernel of the factored CSR operator (csrc/csr_ops.cu) against the launch-chain
solver and a direct solve (mode-solver inner solve, Mode_Solver_2D.py:776-781 `eigs(..., sigma=...)`)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu


def _stencil(rows, cols, offsets, rng):
    """Ragged grid-like CSR factor: row i touches the columns (i*cols//rows + d) % cols, d in offsets; every
    17th row is empty, every 5th loses its last entry.  (sp.random is unusable at 77k x 77k: it samples
    rows*cols without replacement.)"""
    base = (np.arange(rows, dtype=np.int64) * cols) // rows
    ri, ci = [], []
    for k, d in enumerate(offsets):
        keep = np.ones(rows, bool)
        keep[::17] = False
        if k == len(offsets) - 1:
            keep[::5] = False
        ri.append(np.arange(rows)[keep])
        ci.append(((base + d) % cols)[keep])
    ri = np.concatenate(ri)
    ci = np.concatenate(ci)
    v = rng.standard_normal(ri.size) + 1j * rng.standard_normal(ri.size)
    return sp.coo_matrix((v, (ri, ci)), shape=(rows, cols)).tocsr()


def _factors(n, m, seed):
    """P (n x m), Q (m x n): ragged rows, a few empty rows, row counts not multiples of the warp's row group."""
    rng = np.random.default_rng(seed)
    w = max(2, int(np.sqrt(n)))
    return _stencil(n, m, (0, 1, w), rng), _stencil(m, n, (0, m - 1, m - w), rng), rng


@pytest.mark.parametrize("n,m", [(37, 41), (1003, 997), (77201, 77683)])
def test_persistent_bicgstab_vs_chain_and_direct(n, m):
    import torch
    from fdfd_b200.operators import CsrProductOperator
    P, Q, rng = _factors(n, m, 5)
    omega = (P @ Q).tocsr()
    shift = 4.0 * abs(omega).sum(axis=1).max() + 0.5j        # strictly dominant: both solvers converge quickly
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x_ref = spla.spsolve((omega + shift * sp.identity(n)).tocsc(), b) if n <= 5000 else None   # LU fill at 77k: 25 s
    op = CsrProductOperator(P, Q, shift=shift)
    bd = torch.from_numpy(b).cuda()
    out = {}
    for mode in ("persistent", "chain"):
        if mode == "chain":
            os.environ["FDFD_NO_PERSISTENT_BICGSTAB"] = "1"
        try:
            x, info = op.solve(bd, rtol=1e-12, maxit=500)
        finally:
            os.environ.pop("FDFD_NO_PERSISTENT_BICGSTAB", None)
        assert info["converged"]
        r = b - (omega @ x.cpu().numpy() + shift * x.cpu().numpy())
        assert np.linalg.norm(r) <= 2e-12 * np.linalg.norm(b)          # true residual, recomputed on the host
        if x_ref is not None:
            assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-10 * np.linalg.norm(x_ref)
        out[mode] = (x.cpu().numpy(), info)
    assert abs(out["persistent"][1]["iterations"] - out["chain"][1]["iterations"]) <= 2
    assert np.linalg.norm(out["persistent"][0] - out["chain"][0]) <= 1e-10 * np.linalg.norm(out["chain"][0])
    op.close()


def test_persistent_bicgstab_reports_no_convergence():
    import torch
    from fdfd_b200.operators import CsrProductOperator
    P, Q, rng = _factors(2001, 1999, 11)
    omega = (P @ Q).tocsr()
    b = rng.standard_normal(2001) + 0j
    op = CsrProductOperator(P, Q, shift=0.01)                 # indefinite, far from dominant
    x, info = op.solve(torch.from_numpy(b).cuda(), rtol=1e-14, maxit=5, raise_on_noconv=False)
    assert not info["converged"] and info["iterations"] <= 6
    r = b - (omega @ x.cpu().numpy() + 0.01 * x.cpu().numpy())
    assert abs(np.linalg.norm(r) / np.linalg.norm(b) - info["relres"]) <= 1e-6 * max(1.0, info["relres"])
    op.close()
