"""ORACLE (test infrastructure, not product code) for K5, the shifted-Laplacian multigrid
preconditioner of csrc/mg.cu.  This is NOT a restatement of the reference (the reference
solves with SuperLU, Scattering_Solver_2D.py:194-217; it has no preconditioner) — it is a
numpy/scipy restatement of *this repository's own* algorithm, used to check the CUDA kernels
component by component (same coarsening, transfer operators, smoother, cycle), and it is the
prototype the algorithm was designed with.

Algorithm (identical to mg.cu):
  M_l = sx_l*D ca_l D + sy_l*D cb_l D + shift*cd_l on every level, re-discretised;
  coarse ca/cb = mean of the two fine faces on the coarse face, cd = mean of 4 cells,
  wall-face coefficient scaled by delta_l/delta_{l+1}; restriction = mean of 4;
  prolongation = cell-centred bilinear, clamped at the flux-free side, -> 0 at the wall;
  smoother = omega-Jacobi + y-line solves in the x-strips, then x-line solves in the y-strips
  (only strip rows updated in the second half-sweep); coarsest level = `coarse_its` sweeps.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import yee_oracle


def op_matrix(pol, ca, cb, cd, sx, sy):
    Ny, Nx = ca.shape
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([Nx, Ny], [1.0, 1.0], [0, 0])
    a, b, d = sp.diags(ca.ravel()), sp.diags(cb.ravel()), sp.diags(cd.ravel())
    if pol == "TE":
        return (sx * (DHX @ a @ DEX) + sy * (DHY @ b @ DEY) + d).tocsr()
    return (sx * (DEX @ a @ DHX) + sy * (DEY @ b @ DHY) + d).tocsr()


def coarsen(pol, ca, cb, cd, delta):
    ix = 1 if pol == "TE" else 0
    cac = 0.5 * (ca[0::2, ix::2] + ca[1::2, ix::2])
    cbc = 0.5 * (cb[ix::2, 0::2] + cb[ix::2, 1::2])
    cdc = 0.25 * (cd[0::2, 0::2] + cd[1::2, 0::2] + cd[0::2, 1::2] + cd[1::2, 1::2])
    m = delta / ((delta + 0.5) / 2.0)
    if pol == "TE":
        cac[:, -1] *= m
        cbc[-1, :] *= m
    else:
        cac[:, 0] *= m
        cbc[0, :] *= m
    return cac, cbc, cdc


def restrict(r):
    return 0.25 * (r[0::2, 0::2] + r[1::2, 0::2] + r[0::2, 1::2] + r[1::2, 1::2])


def prolong(e, pol, delta_fine):
    wq = delta_fine / (2.0 * ((delta_fine + 0.5) / 2.0))

    def up(a, axis):
        a = np.moveaxis(a, axis, 0)
        lo = np.concatenate([a[:1], a[:-1]], axis=0)
        hi = np.concatenate([a[1:], a[-1:]], axis=0)
        out = np.empty((2 * a.shape[0],) + a.shape[1:], dtype=a.dtype)
        out[0::2] = 0.75 * a + 0.25 * lo
        out[1::2] = 0.75 * a + 0.25 * hi
        if pol == "TE":
            out[-1] = wq * a[-1]
        else:
            out[0] = wq * a[0]
        return np.moveaxis(out, 0, axis)
    return up(up(e, 0), 1)


class MGOracle:
    def __init__(self, pol, ca, cb, cd, sx, sy, shift=(1.0, 0.5), omega=0.8, pre=1, post=1,
                 lines=(0, 0, 0, 0), levels=0, min_n=32, coarse_its=0, cycle=1, wfrom=3, coarse_mode=1,
                 kh_max=1.3):
        self.pol, self.omega, self.pre, self.post = pol, omega, pre, post
        self.coarse_its, self.cycle_kind, self.wfrom = coarse_its, cycle, wfrom
        self.coarse_mode = coarse_mode
        self.levels = []
        sh = shift[0] + 1j * shift[1]
        cdm = sh * np.asarray(cd, dtype=complex)
        ca = np.asarray(ca, dtype=complex)
        cb = np.asarray(cb, dtype=complex)
        delta = 1.0
        lx_lo, lx_hi, ly_lo, ly_hi = lines
        while True:
            Ny, Nx = ca.shape
            A = op_matrix(pol, ca, cb, cdm, sx, sy)
            L = dict(A=A, shape=(Ny, Nx), delta=delta)
            jj, ii = np.meshgrid(np.arange(Ny), np.arange(Nx), indexing="ij")
            wl, wh = (lx_lo, lx_hi) if lx_lo + lx_hi <= Nx else (Nx, 0)
            vl, vh = (ly_lo, ly_hi) if ly_lo + ly_hi <= Ny else (Ny, 0)
            inx = ((ii < wl) | (ii >= Nx - wh)).ravel()
            iny = ((jj < vl) | (jj >= Ny - vh)).ravel()
            C = A.tocoo()
            dcol = C.col - C.row

            def blk(keep):
                return spla.splu(sp.csc_matrix((C.data[keep], (C.row[keep], C.col[keep])), shape=A.shape))
            L["B1"] = blk((dcol == 0) | ((abs(dcol) == Nx) & inx[C.row] & inx[C.col]))
            L["B2"] = blk((dcol == 0) | ((abs(dcol) == 1) & iny[C.row] & iny[C.col])) if iny.any() else None
            L["m2"] = iny.astype(float)
            self.levels.append(L)
            kh_next = 1.0 / np.sqrt(min(sx, sy) / 4.0)
            stop = ((levels and len(self.levels) >= levels) or Nx % 2 or Ny % 2 or Nx // 2 < min_n
                    or Ny // 2 < min_n or kh_next > kh_max)
            L["kh"] = 1.0 / np.sqrt(min(sx, sy))
            if stop:
                break
            ca, cb, cdm = coarsen(pol, ca, cb, cdm, delta)
            delta = (delta + 0.5) / 2.0
            sx, sy = sx / 4, sy / 4
            lx_lo, lx_hi, ly_lo, ly_hi = [(w + 1) // 2 for w in (lx_lo, lx_hi, ly_lo, ly_hi)]

    def _coarse_its(self):
        if self.coarse_its > 0:
            return self.coarse_its
        return int(min(80, max(8, np.ceil(10.0 / self.levels[-1]["kh"]))))

    def _coarse_bicgstab(self, l, x, b, its):
        """`its` iterations of right-preconditioned BiCGSTAB (preconditioner = one smoothing sweep
        from zero), same recurrences as csrc/krylov.cu."""
        A = self.levels[l]["A"]

        def sdiv(a, c):
            return a / c if (abs(c) > 0 and np.isfinite(abs(c))) else 0.0
        M = lambda v: self.smooth(l, np.zeros_like(v), v)  # noqa: E731
        r = b - A @ x
        rh = r.copy()
        rho = np.vdot(r, r)
        rho_old = alpha = om = 1.0
        v = np.zeros_like(b)
        p = np.zeros_like(b)
        for _ in range(its):
            beta = sdiv(rho, rho_old) * sdiv(alpha, om)
            p = r + beta * p - (beta * om) * v
            ph = M(p)
            v = A @ ph
            alpha = sdiv(rho, np.vdot(rh, v))
            s = r - alpha * v
            sh = M(s)
            t = A @ sh
            om = sdiv(np.vdot(t, s), np.vdot(t, t))
            x = x + alpha * ph + om * sh
            rho_old = rho
            r = s - om * t
            rho = np.vdot(rh, r)
        return x

    def smooth(self, l, x, b):
        L = self.levels[l]
        x = x + self.omega * L["B1"].solve(b - L["A"] @ x)
        if L["B2"] is not None:
            x = x + L["m2"] * (self.omega * L["B2"].solve(b - L["A"] @ x))
        return x

    def cycle(self, l, x, b):
        L = self.levels[l]
        last = len(self.levels) - 1
        if l == last:
            if self.coarse_mode == 1:
                return self._coarse_bicgstab(l, x, b, self._coarse_its())
            for _ in range(self._coarse_its()):
                x = self.smooth(l, x, b)
            return x
        for _ in range(self.pre):
            x = self.smooth(l, x, b)
        r = b - L["A"] @ x
        bc = restrict(r.reshape(L["shape"])).ravel()
        xc = np.zeros_like(bc)
        gamma = 2 if ((self.cycle_kind == 2 or (self.wfrom >= 0 and l + 1 >= self.wfrom)) and l + 1 < last) else 1
        for _ in range(gamma):
            xc = self.cycle(l + 1, xc, bc)
        x = x + prolong(xc.reshape(self.levels[l + 1]["shape"]), self.pol, L["delta"]).ravel()
        for _ in range(self.post):
            x = self.smooth(l, x, b)
        return x

    def __call__(self, v):
        v = np.asarray(v, dtype=complex).ravel()
        return self.cycle(0, np.zeros_like(v), v)
