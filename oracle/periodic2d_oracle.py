"""ORACLE (test infrastructure, not product code): CPU restatement of the 2-D Bloch-periodic TE/TM mode solver
of the reference, /root/reference/Periodic_Solver_2D/Periodic_Solver_2D.py:

  staggered shapes, constants            :41-67   (eps0 = 8.85e-12, mu0 = 1.26e-6, c = 1/sqrt(eps0 mu0), k0 = omega/c)
  x difference, periodic z difference
  and forward average                    :198-267 (vectorised here; same CSR arrays)
  cell -> field averaging                :382-412 (eps_xx, mu_yy: periodic z average; eps_yy, eps_zz, mu_xx: x average)
  TM / TE pencils (A, B)                 :484-512
  free mask, eigs(A, M=B, sigma, v0=1)   :514-566

Material pre-processing (sub-pixel rectangles, 1e8-penalty PEC / PMC, x-directed PML) belongs to the product class
and is compared with the reference directly in tests/test_oracle_vs_reference.py.  Pinned there (operators and
pencils bit for bit against the unmodified reference) and by tests/golden/periodic2d_golden.npz.

What the reference's `eigs` call returns is reproducible only for the pairs ARPACK actually converges
(scripts/periodic2d_reproducibility.py): `solve` below therefore also returns the relative residual of every
pair in the pencil, and the parity tests compare converged pairs only.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS0 = 8.85e-12
MU0 = 1.26e-6


def shapes(Nx, Nz):
    return dict(cell=(Nx, Nz), ex=(Nx, Nz), hy=(Nx, Nz), ez=(Nx + 1, Nz), hx=(Nx + 1, Nz), ey=(Nx + 1, Nz), hz=(Nx, Nz))


def _grid(shape):
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    return i.ravel(order="F"), j.ravel(order="F")


def difference_x(in_shape, out_shape, scale):
    """:218-233 (forward): +1/scale at (i+1, j), -1/scale at (i, j); entries outside the input grid dropped."""
    i, j = _grid(out_shape)
    row = i + j * out_shape[0]
    r, c, v = [], [], []
    for di, val in ((1, 1.0), (0, -1.0)):
        ok = (i + di < in_shape[0]) & (j < in_shape[1])
        r.append(row[ok]); c.append((i + di + j * in_shape[0])[ok]); v.append(np.full(int(ok.sum()), val / scale))
    return sp.coo_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))),
                         shape=(int(np.prod(out_shape)), int(np.prod(in_shape)))).tocsr()


def periodic_dz(shape, scale):
    """:235-253: (f[(j+1) % nz] - f[j]) / dz, no Bloch phase."""
    i, j = _grid(shape)
    row = i + j * shape[0]
    col = i + ((j + 1) % shape[1]) * shape[0]
    n = row.size
    return sp.coo_matrix((np.concatenate([np.full(n, 1.0 / scale), np.full(n, -1.0 / scale)]),
                          (np.concatenate([row, row]), np.concatenate([col, row]))), shape=(n, n)).tocsr()


def periodic_az(shape):
    """:255-267: 0.5 (f[j] + f[(j+1) % nz])."""
    i, j = _grid(shape)
    row = i + j * shape[0]
    col = i + ((j + 1) % shape[1]) * shape[0]
    n = row.size
    return sp.coo_matrix((np.full(2 * n, 0.5), (np.concatenate([row, row]), np.concatenate([row, col]))),
                         shape=(n, n)).tocsr()


def operators(Nx, Nz, dx, dz):
    """:198-213."""
    s = shapes(Nx, Nz)
    D = dict(DEX_EZ_TO_EX=difference_x(s["ez"], s["ex"], dx), DEX_EY_TO_HZ=difference_x(s["ey"], s["hz"], dx),
             DEZ_EX=periodic_dz(s["ex"], dz), DEZ_EY=periodic_dz(s["ey"], dz),
             AZ_EX=periodic_az(s["ex"]), AZ_EY=periodic_az(s["ey"]))
    D["DHX_HY_TO_EZ"] = -D["DEX_EZ_TO_EX"].conj().T
    D["DHX_HZ_TO_HX"] = -D["DEX_EY_TO_HZ"].conj().T
    D["DHZ_HY"] = -D["DEZ_EX"].conj().T
    D["DHZ_HX"] = -D["DEZ_EY"].conj().T
    D["AZ_HY"] = D["AZ_EX"].conj().T
    D["AZ_HX"] = D["AZ_EY"].conj().T
    return D


def average_z(values, keep=None):
    """:382-388."""
    out = 0.5 * (values + np.roll(values, -1, axis=1))
    if keep is not None:
        ii, jj = np.nonzero(keep)
        out[ii, jj] = values[ii, jj]
        out[ii, (jj - 1) % values.shape[1]] = values[ii, jj]
    return out


def average_x(values, keep=None):
    """:390-402."""
    nx, nz = values.shape
    acc = np.zeros((nx + 1, nz), dtype=complex)
    cnt = np.zeros((nx + 1, nz), dtype=float)
    acc[:nx] += values; cnt[:nx] += 1
    acc[1:] += values; cnt[1:] += 1
    acc = acc / cnt
    if keep is not None:
        ii, jj = np.nonzero(keep)
        acc[ii, jj] = values[ii, jj]
        acc[ii + 1, jj] = values[ii, jj]
    return acc


def materials_on_fields(cell, keep):
    """:404-412.  `cell`: dict eps_xx ... mu_zz of (Nx, Nz) arrays."""
    return dict(eps_xx=average_z(cell["eps_xx"], keep), eps_yy=average_x(cell["eps_yy"], keep),
                eps_zz=average_x(cell["eps_zz"], keep), mu_xx=average_x(cell["mu_xx"], keep),
                mu_yy=average_z(cell["mu_yy"], keep), mu_zz=cell["mu_zz"].copy())


def _dg(a):
    return sp.diags(a.ravel(order="F"), format="csr")


def pencil(pol, mats, D, omega):
    """:484-512."""
    if pol == "TM":
        n = D["DEZ_EX"].shape[0]
        zero = sp.csr_matrix((n, n), dtype=complex)
        D1 = 1j * omega * MU0 * _dg(mats["mu_yy"])
        D1 += 1j / omega * D["DEX_EZ_TO_EX"] @ (1 / EPS0 * _dg(mats["eps_zz"]).power(-1) @ D["DHX_HY_TO_EZ"])
        D2 = 1j * omega * EPS0 * _dg(mats["eps_xx"])
        A = sp.bmat([[D["DEZ_EX"], D1], [D2, D["DHZ_HY"]]], format="csr")
        B = sp.bmat([[D["AZ_EX"], zero], [zero, D["AZ_HY"]]], format="csr")
    else:
        n = D["DHZ_HX"].shape[0]
        zero = sp.csr_matrix((n, n), dtype=complex)
        D1 = -1j * omega * EPS0 * _dg(mats["eps_yy"])
        D1 -= 1j / omega * D["DHX_HZ_TO_HX"] @ (1 / MU0 * _dg(mats["mu_zz"]).power(-1) @ D["DEX_EY_TO_HZ"])
        D2 = -1j * omega * MU0 * _dg(mats["mu_xx"])
        A = sp.bmat([[D["DHZ_HX"], D1], [D2, D["DEZ_EY"]]], format="csr")
        B = sp.bmat([[D["AZ_HX"], zero], [zero, D["AZ_EY"]]], format="csr")
    return A, B


def relative_residuals(A, B, lam, X):
    """||A x - lambda B x|| / (||A x|| + |lambda| ||B x||) per column."""
    out = []
    for i in range(len(lam)):
        ax, bx = A @ X[:, i], B @ X[:, i]
        den = np.linalg.norm(ax) + abs(lam[i]) * np.linalg.norm(bx)
        out.append(float(np.linalg.norm(ax - lam[i] * bx) / den) if den > 0 else 0.0)
    return np.array(out)


def solve(pol, mats, Nx, Nz, dx, dz, omega, num_modes, guess, free=None, tol=0, ncv=None):
    """:523-566: eigs(A, M=B, k, sigma=guess, v0=ones) on the free unknowns, ordered by |lambda - guess|.
    Returns (eigenvalues, reduced eigenvectors, relative residuals)."""
    A, B = pencil(pol, mats, operators(Nx, Nz, dx, dz), omega)
    if free is not None:
        A, B = A[free, :][:, free], B[free, :][:, free]
    lam, X = spla.eigs(A, M=B, k=num_modes, sigma=guess, tol=tol, ncv=ncv, v0=np.ones(A.shape[0], dtype=complex))
    order = np.argsort(np.abs(lam - guess))
    lam, X = lam[order], X[:, order]
    return lam, X, relative_residuals(A, B, lam, X)
