"""ORACLE (test infrastructure, not product code): CPU restatement of the full-vector 2-D mode
solver hot path of the reference, /root/reference/Mode_Solver_2D/Mode_Solver_2D.py, from the
cell-centred material tensors and PEC/PMC masks to eigenvalues, n_eff and fields:

  staggered shapes                    Mode_Solver_2D.py:33-48
  rectangular difference operators    :482-544   (vectorised here; same CSR arrays)
  cell -> Yee-location averaging      :546-602
  effective materials and masks       :626-676
  P, Q, Omega = P_red Q_red           :734-774
  eigs(Omega, k, sigma)               :780       (ARPACK mode 3 + SuperLU, scipy)
  n_eff branch rules                  :830-840
  H, Ez, Hz reconstruction            :792-805

Geometry helpers (sub-pixel fill fractions, UPML scaling of the cell tensors, PEC regions) are
pre-processing: the product class implements them and tests compare its cell arrays with the
reference's directly.  Pinned by tests/test_oracle_vs_reference.py::test_mode_* (reference imported
in-container) and tests/golden/mode_golden.npz (SURVEY.md §8c eigenvalues of config 2).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigs


def shapes(Nx, Ny):
    return dict(cell=(Nx, Ny), ex=(Nx, Ny + 1), ey=(Nx + 1, Ny), ez=(Nx + 1, Ny + 1),
                hx=(Nx + 1, Ny), hy=(Nx, Ny + 1), hz=(Nx, Ny))


def difference(in_shape, out_shape, scale, axis):
    """Forward difference between two staggered grids (Mode_Solver_2D.py:482-514): row (i, j) of the
    output grid holds +1/scale at the forward neighbour and -1/scale at (i, j) of the input grid
    (flat index i + j*nx, Fortran order), entries outside the input grid dropped."""
    onx, ony = out_shape
    inx, iny = in_shape
    i, j = np.meshgrid(np.arange(onx), np.arange(ony), indexing="ij")
    i = i.ravel(order="F")
    j = j.ravel(order="F")
    row = i + j * onx
    rows, cols, vals = [], [], []
    for di, dj, v in (((1, 0) if axis == 0 else (0, 1)) + (1.0,), (0, 0, -1.0)):
        ci, cj = i + di, j + dj
        ok = (ci >= 0) & (ci < inx) & (cj >= 0) & (cj < iny)
        rows.append(row[ok])
        cols.append((ci + cj * inx)[ok])
        vals.append(np.full(ok.sum(), v / scale))
    return sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(onx * ony, inx * iny)).tocsr()


def operators(Nx, Ny, dxn, dyn):
    """The eight operators of `_yeeder2d` (Mode_Solver_2D.py:516-544)."""
    s = shapes(Nx, Ny)
    d_ez_ex = difference(s["ez"], s["ex"], dxn, 0)
    d_ez_ey = difference(s["ez"], s["ey"], dyn, 1)
    d_ey_hz = difference(s["ey"], s["hz"], dxn, 0)
    d_ex_hz = difference(s["ex"], s["hz"], dyn, 1)
    return dict(d_ez_ex=d_ez_ex, d_ez_ey=d_ez_ey, d_ey_hz=d_ey_hz, d_ex_hz=d_ex_hz,
                d_hy_ez=-d_ez_ex.conj().T, d_hx_ez=-d_ez_ey.conj().T,
                d_hz_hx=-d_ey_hz.conj().T, d_hz_hy=-d_ex_hz.conj().T)


def _avg(values, shape, along):
    """Mean of the adjacent cells of each Yee location (fewer at the domain boundary)."""
    out = np.zeros(shape, dtype=complex)
    cnt = np.zeros(shape, dtype=float)
    Nx, Ny = values.shape
    offs = {"y": ((0, 0), (0, 1)), "x": ((0, 0), (1, 0)), "xy": ((0, 0), (1, 0), (0, 1), (1, 1))}[along]
    for ox, oy in offs:
        out[ox:ox + Nx, oy:oy + Ny] += values
        cnt[ox:ox + Nx, oy:oy + Ny] += 1
    return out / cnt, offs


def average_to(values, shape, along, no_average_mask=None):
    out, offs = _avg(values, shape, along)
    if no_average_mask is not None:
        ii, jj = np.nonzero(no_average_mask)
        for ox, oy in offs:
            out[ii + ox, jj + oy] = values[ii, jj]
    return out


def materials_on_fields(cell, no_average_mask, Nx, Ny):
    """Mode_Solver_2D.py:604-611.  `cell` maps eps_xx.. mu_zz to (Nx, Ny) arrays."""
    s = shapes(Nx, Ny)
    return dict(eps_xx=average_to(cell["eps_xx"], s["ex"], "y", no_average_mask),
                eps_yy=average_to(cell["eps_yy"], s["ey"], "x", no_average_mask),
                eps_zz=average_to(cell["eps_zz"], s["ez"], "xy", no_average_mask),
                mu_xx=average_to(cell["mu_xx"], s["ey"], "x", no_average_mask),
                mu_yy=average_to(cell["mu_yy"], s["ex"], "y", no_average_mask),
                mu_zz=cell["mu_zz"].copy())


def masks_from_cells(cell_mask, Nx, Ny, field):
    """Mode_Solver_2D.py:369-402."""
    s = shapes(Nx, Ny)
    ii, jj = np.nonzero(cell_mask)
    if field == "electric":
        xx, yy, zz = (np.zeros(s[k], bool) for k in ("ex", "ey", "ez"))
        for o in (0, 1):
            xx[ii, jj + o] = True
            yy[ii + o, jj] = True
            for p in (0, 1):
                zz[ii + o, jj + p] = True
    else:
        xx, yy, zz = (np.zeros(s[k], bool) for k in ("hx", "hy", "hz"))
        for o in (0, 1):
            xx[ii + o, jj] = True
            yy[ii, jj + o] = True
        zz[ii, jj] = True
    return xx, yy, zz


def effective(cell, no_average_mask, pec, pmc, Nx, Ny):
    """Mode_Solver_2D.py:626-676: non-finite cell values become PEC/PMC masks, constrained material
    entries are set to 1.  `pec`/`pmc` map 'xx','yy','zz' to boolean masks (copied)."""
    cell = {k: v.copy() for k, v in cell.items()}
    pec = {k: v.copy() for k, v in pec.items()}
    pmc = {k: v.copy() for k, v in pmc.items()}
    idx = {"xx": 0, "yy": 1, "zz": 2}
    for comp in ("xx", "yy", "zz"):
        for pre, target, fld in (("eps_", pec, "electric"), ("mu_", pmc, "magnetic")):
            vals = cell[pre + comp]
            bad = ~np.isfinite(vals)
            if np.any(bad):
                target[comp] |= masks_from_cells(bad, Nx, Ny, fld)[idx[comp]]
                vals[bad] = 1.0 + 0j
    mats = materials_on_fields(cell, no_average_mask, Nx, Ny)
    for comp in ("xx", "yy", "zz"):
        mats["eps_" + comp][pec[comp]] = 1.0 + 0j
        mats["mu_" + comp][pmc[comp]] = 1.0 + 0j
    return mats, pec, pmc


def assemble(mats, pec, pmc, Nx, Ny, dxn, dyn):
    """P, Q (full), the free-DOF masks and the reduced factors (Mode_Solver_2D.py:734-774)."""
    D = operators(Nx, Ny, dxn, dyn)
    diag = lambda a: sp.diags(a.ravel(order="F"), format="csr")  # noqa: E731

    def inv_free(values, constrained):
        flat = values.ravel(order="F")
        c = constrained.ravel(order="F")
        inv = np.zeros_like(flat, dtype=complex)
        inv[~c] = 1.0 / flat[~c]
        return sp.diags(inv, format="csr")
    ezi = inv_free(mats["eps_zz"], pec["zz"])
    mzi = inv_free(mats["mu_zz"], pmc["zz"])
    P = sp.bmat([[D["d_ez_ex"] @ ezi @ D["d_hx_ez"], -(D["d_ez_ex"] @ ezi @ D["d_hy_ez"] + diag(mats["mu_yy"]))],
                 [D["d_ez_ey"] @ ezi @ D["d_hx_ez"] + diag(mats["mu_xx"]), -D["d_ez_ey"] @ ezi @ D["d_hy_ez"]]],
                format="csr")
    Q = sp.bmat([[D["d_hz_hx"] @ mzi @ D["d_ex_hz"], -(D["d_hz_hx"] @ mzi @ D["d_ey_hz"] + diag(mats["eps_yy"]))],
                 [D["d_hz_hy"] @ mzi @ D["d_ex_hz"] + diag(mats["eps_xx"]), -D["d_hz_hy"] @ mzi @ D["d_ey_hz"]]],
                format="csr")
    free_e = np.concatenate((~pec["xx"].ravel(order="F"), ~pec["yy"].ravel(order="F")))
    free_h = np.concatenate((~pmc["xx"].ravel(order="F"), ~pmc["yy"].ravel(order="F")))
    P_red = P[:, free_h]
    Q_red = Q[free_h, :]
    return dict(P=P, Q=Q, P_red=P_red, Q_red=Q_red, free_e=free_e, free_h=free_h, D=D, ezi=ezi, mzi=mzi)


def passive_neff(neff_squared, lossy):
    """Mode_Solver_2D.py:830-840."""
    root = np.sqrt(neff_squared)
    neff = np.where(np.real(root) < 0, -root, root)
    re, im = np.real(neff), np.imag(neff)
    tol = 1e-12 * np.maximum(1.0, np.abs(neff))
    re = np.where(np.abs(re) <= tol, 0.0, re)
    im = np.where(np.abs(im) <= tol, 0.0, np.abs(im))
    if not lossy:
        im = np.zeros_like(im)
    return re + 1j * im


def solve(cell, no_average_mask, pec, pmc, Nx, Ny, dxn, dyn, num_modes, sigma, lossy):
    """eigenvalues (ascending real part), n_eff, reduced system pieces."""
    mats, pec, pmc = effective(cell, no_average_mask, pec, pmc, Nx, Ny)
    sysm = assemble(mats, pec, pmc, Nx, Ny, dxn, dyn)
    Omega = (sysm["P_red"] @ sysm["Q_red"])[sysm["free_e"], :][:, sysm["free_e"]]
    vals, vecs = eigs(Omega, k=num_modes, sigma=sigma)
    order = np.argsort(np.real(vals))
    vals = vals[order]
    full = np.zeros((sysm["free_e"].size, num_modes), complex)
    full[sysm["free_e"], :] = vecs[:, order]
    return dict(eigenvalues=vals, neff=passive_neff(-vals, lossy), eigenvectors=full, Omega=Omega, sys=sysm, mats=mats)
