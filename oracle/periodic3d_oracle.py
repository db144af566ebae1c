"""ORACLE (test infrastructure, not product code): CPU restatement of the 3-D periodic mode
solver hot path of the reference,
  /root/reference/Periodic_Solver_3D/Periodic_Solver_3D.py and
  /root/reference/Periodic_Solver_3D/refined_shift_invert_arnoldi.py:

  staggered shapes, constants           Periodic_Solver_3D.py:18-59 (k0 = omega/3e8, eps0 = 8.85e-12, mu0 = 1.26e-6)
  x / y differences, periodic z
  difference and forward average        :95-194  (vectorised here; same CSR arrays)
  cell -> Yee averaging                 :398-475
  4x4 block pencil (A, B)               :478-533
  refined shift-invert Arnoldi          refined_shift_invert_arnoldi.py:5-164
      (MGS + second pass, eig(H), theta = sigma + 1/mu sorted by |theta - sigma|, per-candidate
       SVD of (A - theta B) V -> refined vector, relative residual, restart from the sum of
       refined vectors + 1e-3 noise)

Material pre-processing (sub-pixel blocks, 1e8-penalty PEC/PMC, UPML division/multiplication per
layer) belongs to the product class and is compared with the reference directly in
tests/test_oracle_vs_reference.py.  Pinned there and by tests/golden/periodic3d_golden.npz.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import splu


def shapes(Nx, Ny, Nz):
    return dict(cell=(Nx, Ny, Nz), ex=(Nx, Ny + 1, Nz), ey=(Nx + 1, Ny, Nz), ez=(Nx + 1, Ny + 1, Nz),
                hx=(Nx + 1, Ny, Nz), hy=(Nx, Ny + 1, Nz), hz=(Nx, Ny, Nz))


def _grid(shape):
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    return i.ravel(order="F"), j.ravel(order="F"), k.ravel(order="F")


def _flat(i, j, k, nx, ny):
    return i + nx * (j + ny * k)


def difference_xy(in_shape, out_shape, scale, axis):
    """Periodic_Solver_3D.py:120-158 (forward): +1/scale at the forward neighbour, -1/scale at self,
    entries outside the input grid dropped."""
    i, j, k = _grid(out_shape)
    row = _flat(i, j, k, out_shape[0], out_shape[1])
    r, c, v = [], [], []
    for di, dj, val in (((1, 0) if axis == 0 else (0, 1)) + (1.0,), (0, 0, -1.0)):
        ci, cj = i + di, j + dj
        ok = (ci < in_shape[0]) & (cj < in_shape[1]) & (k < in_shape[2])
        r.append(row[ok]); c.append(_flat(ci, cj, k, in_shape[0], in_shape[1])[ok]); v.append(np.full(int(ok.sum()), val / scale))
    return sp.coo_matrix((np.concatenate(v), (np.concatenate(r), np.concatenate(c))),
                         shape=(int(np.prod(out_shape)), int(np.prod(in_shape)))).tocsr()


def periodic_dz(shape, scale):
    """Periodic_Solver_3D.py:160-179: (f_{k+1 mod nz} - f_k) / dz, no Bloch phase."""
    i, j, k = _grid(shape)
    row = _flat(i, j, k, shape[0], shape[1])
    col_f = _flat(i, j, (k + 1) % shape[2], shape[0], shape[1])
    n = row.size
    return sp.coo_matrix((np.concatenate([np.full(n, 1.0 / scale), np.full(n, -1.0 / scale)]),
                          (np.concatenate([row, row]), np.concatenate([col_f, row]))), shape=(n, n)).tocsr()


def periodic_az(shape):
    """Periodic_Solver_3D.py:181-194: 0.5 (f_k + f_{k+1 mod nz})."""
    i, j, k = _grid(shape)
    row = _flat(i, j, k, shape[0], shape[1])
    col_f = _flat(i, j, (k + 1) % shape[2], shape[0], shape[1])
    n = row.size
    return sp.coo_matrix((np.full(2 * n, 0.5), (np.concatenate([row, row]), np.concatenate([row, col_f]))),
                         shape=(n, n)).tocsr()


def operators(Nx, Ny, Nz, dx, dy, dz):
    s = shapes(Nx, Ny, Nz)
    D = dict(DEX_EZ_TO_EX=difference_xy(s["ez"], s["ex"], dx, 0), DEY_EZ_TO_EY=difference_xy(s["ez"], s["ey"], dy, 1),
             DEX_EY_TO_HZ=difference_xy(s["ey"], s["hz"], dx, 0), DEY_EX_TO_HZ=difference_xy(s["ex"], s["hz"], dy, 1),
             DEZ_EX=periodic_dz(s["ex"], dz), DEZ_EY=periodic_dz(s["ey"], dz),
             AZ_EX=periodic_az(s["ex"]), AZ_EY=periodic_az(s["ey"]))
    D["DHX_HY_TO_EZ"] = -D["DEX_EZ_TO_EX"].conj().T
    D["DHY_HX_TO_EZ"] = -D["DEY_EZ_TO_EY"].conj().T
    D["DHX_HZ_TO_HX"] = -D["DEX_EY_TO_HZ"].conj().T
    D["DHY_HZ_TO_HY"] = -D["DEY_EX_TO_HZ"].conj().T
    D["DHZ_HX"] = -D["DEZ_EY"].conj().T
    D["DHZ_HY"] = -D["DEZ_EX"].conj().T
    D["AZ_HX"] = D["AZ_EY"].conj().T
    D["AZ_HY"] = D["AZ_EX"].conj().T
    return D


def average(values, along, keep=None):
    """Periodic_Solver_3D.py:398-454 (x: Nx+1, y: Ny+1, xy: both)."""
    Nx, Ny, Nz = values.shape
    offs = {"x": ((0, 0), (1, 0)), "y": ((0, 0), (0, 1)), "xy": ((0, 0), (1, 0), (0, 1), (1, 1))}[along]
    shape = (Nx + max(o[0] for o in offs), Ny + max(o[1] for o in offs), Nz)
    acc = np.zeros(shape, dtype=complex)
    cnt = np.zeros(shape, dtype=float)
    for ox, oy in offs:
        acc[ox:ox + Nx, oy:oy + Ny, :] += values
        cnt[ox:ox + Nx, oy:oy + Ny, :] += 1
    acc = acc / cnt
    if keep is not None:
        ii, jj, kk = np.nonzero(keep)
        for ox, oy in offs:
            acc[ii + ox, jj + oy, kk] = values[ii, jj, kk]
    return acc


def materials_on_fields(cell, keep):
    """Periodic_Solver_3D.py:456-464.  cell: dict erxx..mrzz of (Nx, Ny, Nz) arrays."""
    return dict(erxx=average(cell["erxx"], "y", keep), eryy=average(cell["eryy"], "x", keep),
                erzz=average(cell["erzz"], "xy", keep), mrxx=average(cell["mrxx"], "x", keep),
                mryy=average(cell["mryy"], "y", keep), mrzz=cell["mrzz"].copy())


def pencil(mats, D, omega, eps0=8.85e-12, mu0=1.26e-6):
    """A, B of Periodic_Solver_3D.py:478-533 (unknowns [Ex; Ey; Hx; Hy])."""
    dg = lambda a: sp.diags(a.ravel(order="F"), format="csr")  # noqa: E731
    Erxx, Eryy, Erzz, Mrxx, Mryy, Mrzz = (dg(mats[k]) for k in ("erxx", "eryy", "erzz", "mrxx", "mryy", "mrzz"))
    ce, cm = 1j / (omega * eps0), 1j / (omega * mu0)
    ezi, mzi = Erzz.power(-1), Mrzz.power(-1)
    n_ex, n_ey = D["DEZ_EX"].shape[0], D["DEZ_EY"].shape[0]
    Z = lambda r, c: sp.csr_matrix((r, c), dtype=complex)  # noqa: E731
    A = sp.bmat([
        [D["DEZ_EX"], Z(n_ex, n_ey), D["DEX_EZ_TO_EX"] @ (-ce * ezi @ D["DHY_HX_TO_EZ"]),
         D["DEX_EZ_TO_EX"] @ (ce * ezi @ D["DHX_HY_TO_EZ"]) + 1j * omega * mu0 * Mryy],
        [Z(n_ey, n_ex), D["DEZ_EY"], D["DEY_EZ_TO_EY"] @ (-ce * ezi @ D["DHY_HX_TO_EZ"]) - 1j * omega * mu0 * Mrxx,
         D["DEY_EZ_TO_EY"] @ (ce * ezi @ D["DHX_HY_TO_EZ"])],
        [D["DHX_HZ_TO_HX"] @ (cm * mzi @ D["DEY_EX_TO_HZ"]),
         D["DHX_HZ_TO_HX"] @ (-cm * mzi @ D["DEX_EY_TO_HZ"]) - 1j * omega * eps0 * Eryy, D["DHZ_HX"], Z(n_ey, n_ex)],
        [D["DHY_HZ_TO_HY"] @ (cm * mzi @ D["DEY_EX_TO_HZ"]) + 1j * omega * eps0 * Erxx,
         D["DHY_HZ_TO_HY"] @ (-cm * mzi @ D["DEX_EY_TO_HZ"]), Z(n_ex, n_ey), D["DHZ_HY"]],
    ]).tocsr()
    B = sp.bmat([[D["AZ_EX"], Z(n_ex, n_ey), None, None], [Z(n_ey, n_ex), D["AZ_EY"], None, None],
                 [None, None, D["AZ_HX"], Z(n_ey, n_ex)], [None, None, Z(n_ex, n_ey), D["AZ_HY"]]], format="csr")
    return A, B


# ---- refined shift-invert Arnoldi (refined_shift_invert_arnoldi.py) ---------------------------------
def default_ncv(n, k, ncv):
    if n <= k + 1:
        raise ValueError(f"Not enough DOFs to solve {k} modes.")
    ncv = max(20, 4 * int(k) + 8) if ncv is None else int(ncv)
    return min(n - 1, max(k + 2, ncv))


def arnoldi(op, n, ncv, v0):
    V = np.zeros((n, ncv + 1), complex)
    H = np.zeros((ncv + 1, ncv), complex)
    V[:, 0] = v0 / np.linalg.norm(v0)
    count = 1
    for j in range(ncv):
        w = op(V[:, j])
        for _ in range(2):                       # MGS + the reference's second pass
            for i in range(j + 1):
                h = np.vdot(V[:, i], w)
                H[i, j] += h
                w = w - h * V[:, i]
        H[j + 1, j] = np.linalg.norm(w)
        if H[j + 1, j] <= 1e-14:
            count = j + 1
            break
        if j + 1 < ncv:
            V[:, j + 1] = w / H[j + 1, j]
            count = j + 2
    return V[:, :count], H[:count, :count]


def refined_candidates(A, B, V, H, sigma, k):
    if H.shape[0] < k:
        raise RuntimeError("Arnoldi subspace is smaller than the requested number of modes.")
    mu = np.linalg.eig(H)[0]
    mu = mu[np.isfinite(mu) & (np.abs(mu) > 1e-14)]
    if mu.size < k:
        raise RuntimeError("Not enough finite Ritz values were generated.")
    theta = sigma + 1.0 / mu
    AV, BV = A @ V, B @ V
    found, seen = [], []
    for t in theta[np.argsort(np.abs(theta - sigma))]:
        if any(abs(t - s) <= 1e-8 * max(1.0, abs(t)) for s in seen):
            continue
        y = np.linalg.svd(AV - t * BV, full_matrices=False)[2].conj().T[:, -1]
        x = V @ y
        nx = np.linalg.norm(x)
        if nx == 0:
            continue
        x = x / nx
        Ax, Bx = A @ x, B @ x
        scale = np.linalg.norm(Ax) + abs(t) * np.linalg.norm(Bx)
        found.append((np.linalg.norm(Ax - t * Bx) / (scale if scale else 1.0), t, x))
        seen.append(t)
    if len(found) < k:
        raise RuntimeError("Refined extraction produced too few candidate modes.")
    sel = sorted(found, key=lambda c: abs(c[1] - sigma))[:k]
    return (np.array([c[1] for c in sel], complex), np.column_stack([c[2] for c in sel]),
            np.array([c[0] for c in sel], float))


def refined_shift_invert_arnoldi(A, B, sigma, k, tol, ncv=None, max_restarts=12, seed=0):
    n = A.shape[0]
    ncv = default_ncv(n, k, ncv)
    conv = max(0.0 if tol is None else float(tol), 1e-12)
    lu = splu((A - sigma * B).tocsc())
    op = lambda v: lu.solve(B @ v)  # noqa: E731
    rng = np.random.default_rng(seed)
    v0 = np.ones(n, complex)
    v0 += 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    best = None
    for restart in range(int(max_restarts) + 1):
        V, H = arnoldi(op, n, ncv, v0)
        vals, vecs, res = refined_candidates(A, B, V, H, sigma, k)
        if best is None or np.max(res) < np.max(best[2]):
            best = (vals, vecs, res, restart)
        if np.all(res <= conv):
            return vals, vecs, res, restart
        v0 = vecs @ np.ones(vecs.shape[1], complex)
        v0 += 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        v0 = v0 / np.linalg.norm(v0)
    return best
