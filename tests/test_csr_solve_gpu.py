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


_MODES = {"resident": {}, "general": {"FDFD_NO_RESIDENT_BICGSTAB": "1"}, "chain": {"FDFD_NO_PERSISTENT_BICGSTAB": "1"}}


def _solve_in_mode(op, mode, b, **kw):
    for k, v in _MODES[mode].items():
        os.environ[k] = v
    try:
        return op.solve(b, **kw)
    finally:
        for k in _MODES[mode]:
            os.environ.pop(k, None)


@pytest.mark.parametrize("n,m", [(37, 41), (1003, 997), (9472, 9473), (77201, 77683), (100003, 99991)])
def test_persistent_bicgstab_vs_chain_and_direct(n, m):
    """The three BiCGSTAB implementations of the factored operator — resident persistent kernel (one row per
    thread, rows in shared memory; up to 148 x 576 rows), general persistent kernel, launch chain — against
    each other, the host-side residual and (small cases) a direct solve."""
    import torch
    from fdfd_b200.operators import CsrProductOperator
    P, Q, rng = _factors(n, m, 5)
    omega = (P @ Q).tocsr()
    shift = 4.0 * abs(omega).sum(axis=1).max() + 0.5j        # strictly dominant: all solvers converge quickly
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x_ref = spla.spsolve((omega + shift * sp.identity(n)).tocsc(), b) if n <= 10000 else None   # LU fill at 77k: 25 s
    op = CsrProductOperator(P, Q, shift=shift)
    bd = torch.from_numpy(b).cuda()
    out = {}
    for mode in _MODES:
        x, info = _solve_in_mode(op, mode, bd, rtol=1e-12, maxit=500)
        assert info["converged"]
        r = b - (omega @ x.cpu().numpy() + shift * x.cpu().numpy())
        assert np.linalg.norm(r) <= 2e-12 * np.linalg.norm(b)          # true residual, recomputed on the host
        if x_ref is not None:
            assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-10 * np.linalg.norm(x_ref)
        out[mode] = (x.cpu().numpy(), info)
    for mode in ("resident", "general"):
        assert abs(out[mode][1]["iterations"] - out["chain"][1]["iterations"]) <= 2
        assert np.linalg.norm(out[mode][0] - out["chain"][0]) <= 1e-10 * np.linalg.norm(out["chain"][0])
    op.close()


@pytest.mark.parametrize("mode", ["resident", "general", "chain"])
def test_bicgstab_modes_on_an_ill_conditioned_system(mode):
    """A weakly dominant shift: well over a hundred iterations with an irregular residual history, so that the
    true-residual confirmation is exercised."""
    import torch
    from fdfd_b200.operators import CsrProductOperator
    n, m = 5003, 4999
    P, Q, rng = _factors(n, m, 21)
    omega = (P @ Q).tocsr()
    shift = 0.155 * abs(omega).sum(axis=1).max() * (1 + 0.3j)    # ~170 Jacobi-BiCGSTAB iterations (scipy: 164-173)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x_ref = spla.spsolve((omega + shift * sp.identity(n)).tocsc(), b)
    op = CsrProductOperator(P, Q, shift=shift)
    x, info = _solve_in_mode(op, mode, torch.from_numpy(b).cuda(), rtol=1e-11, maxit=20000, raise_on_noconv=False)
    r = b - (omega @ x.cpu().numpy() + shift * x.cpu().numpy())
    assert abs(np.linalg.norm(r) / np.linalg.norm(b) - info["relres"]) <= 1e-3 * info["relres"] + 1e-15
    assert info["relres"] <= 1e-9 and 50 <= info["iterations"] <= 2000     # converged or stopped at the rounding floor
    assert np.linalg.norm(x.cpu().numpy() - x_ref) <= 1e-6 * np.linalg.norm(x_ref)
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
