"""ORACLE (test infrastructure, not product code): CPU restatement of the band-diagram hot path
of the reference, /root/reference/Band_Diagram_Solver/Band_Diagram_Solver.py.

  helper grid / materials      Band_Diagram_Solver.py:78-114, :119-170
  Gamma-X-M-Gamma path         :183-251
  Yee tensors from the 2x grid :536-550
  per-beta operators + eigs    :286-323  (A_tm = -DHX URyy^-1 DEX - DHY URxx^-1 DEY, M = ERzz;
                                           A_te = -DEX ERyy^-1 DHX - DEY ERxx^-1 DHY, M = URzz;
                                           scipy eigs(A, M=M, k, sigma) = ARPACK mode 3 + SuperLU)
  sorting / normalisation      :552-560

Eigenvalues are computed with the same scipy calls the reference makes (ARPACK with a random
start vector: eigenvalues are reproducible to ~1e-12, eigenvectors only up to phase).
Pinned by tests/test_oracle_vs_reference.py::test_band_* and tests/golden/band_*.npz.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import eigs

from . import yee_oracle


class BandProblem:
    def __init__(self, a, Nx, Ny=None, b=None, background_er=1.0, background_ur=1.0, bc=(1, 1)):
        self.a = float(a)
        self.b = float(b) if b is not None else float(a)
        self.Nx = int(Nx)
        self.Ny = int(Ny) if Ny is not None else int(Nx)
        self.bc = tuple(bc)
        self.dx, self.dy = self.a / self.Nx, self.b / self.Ny
        xa2 = np.arange(1, 2 * self.Nx + 1) * (self.dx / 2)
        ya2 = np.arange(1, 2 * self.Ny + 1) * (self.dy / 2)
        self.X2, self.Y2 = np.meshgrid(xa2 - np.mean(xa2), ya2 - np.mean(ya2), indexing="ij")
        self.ER2 = np.full((2 * self.Nx, 2 * self.Ny), background_er, dtype=complex)
        self.UR2 = np.full((2 * self.Nx, 2 * self.Ny), background_ur, dtype=complex)

    def add_object(self, mask, er=None, ur=None):
        sel = np.asarray(mask(self.X2, self.Y2) if callable(mask) else mask, dtype=bool)
        if er is not None:
            self.ER2 = np.where(sel, np.asarray(er, dtype=complex), self.ER2)
        if ur is not None:
            self.UR2 = np.where(sel, np.asarray(ur, dtype=complex), self.UR2)

    def add_circle(self, radius, center=(0.0, 0.0), er=None, ur=None):
        self.add_object((self.X2 - center[0]) ** 2 + (self.Y2 - center[1]) ** 2 <= radius ** 2, er=er, ur=ur)

    def tensors(self):
        """Band_Diagram_Solver.py:536-550."""
        return dict(ERxx=self.ER2[1::2, ::2], ERyy=self.ER2[::2, 1::2], ERzz=self.ER2[::2, 1::2],
                    URxx=self.UR2[1::2, ::2], URyy=self.UR2[::2, 1::2], URzz=self.UR2[::2, 1::2])


def gamma_x_m_gamma(a, b):
    gx, gy = 2 * np.pi / a, 2 * np.pi / b
    return [np.array([0.0, 0.0]), 0.5 * np.array([gx, 0.0]), 0.5 * np.array([gx, gy]), np.array([0.0, 0.0])]


def bloch_path(points, total_points):
    """Band_Diagram_Solver.py:197-251."""
    pts = np.asarray(points, dtype=float)
    seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
    total = seg.sum()
    if total == 0:
        seg = np.ones_like(seg)
        total = seg.sum()
    remaining = total_points - 1
    counts = [max(1, int(round(remaining * (length / total)))) for length in seg]
    diff = remaining - sum(counts)
    idx = 0
    while diff != 0 and counts:
        if diff > 0:
            counts[idx % len(counts)] += 1
            diff -= 1
        elif counts[idx % len(counts)] > 1:
            counts[idx % len(counts)] -= 1
            diff += 1
        idx += 1
    betas = [pts[0]]
    ticks = [0]
    acc = 0
    for start, stop, n_seg in zip(pts[:-1], pts[1:], counts):
        for step in range(1, n_seg + 1):
            betas.append(start + (step / n_seg) * (stop - start))
        acc += n_seg
        ticks.append(acc)
    return np.column_stack(betas), ticks


def pencil(p: BandProblem, beta, pol):
    """(A, M) exactly as the reference builds them for one Bloch vector."""
    t = p.tensors()
    diag = lambda a: sp.diags(a.flatten(order="F"))  # noqa: E731
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([p.Nx, p.Ny], [p.dx, p.dy], list(p.bc), beta)
    if pol.upper() == "TM":
        A = -DHX @ diag(t["URyy"]).power(-1) @ DEX - DHY @ diag(t["URxx"]).power(-1) @ DEY
        return A, diag(t["ERzz"])
    A = -DEX @ diag(t["ERyy"]).power(-1) @ DHX - DEY @ diag(t["ERxx"]).power(-1) @ DHY
    return A, diag(t["URzz"])


def bands_at(p: BandProblem, beta, pol, num_bands, sigma=0.0):
    """Sorted eigenvalues (Band_Diagram_Solver.py:313-316 / :320-323) and a/lambda frequencies."""
    A, M = pencil(p, beta, pol)
    vals = eigs(A, M=M, k=num_bands, sigma=sigma)[0]
    vals = vals[np.argsort(np.real(vals))][:num_bands]
    freq = p.a / (2 * np.pi) * np.sqrt(np.clip(np.real_if_close(vals).real, 0.0, None))
    return vals, freq


def band_structure(p: BandProblem, beta_path, num_bands, pols=("TE", "TM"), sigma=0.0):
    out = {pol: (np.zeros((num_bands, beta_path.shape[1]), complex), np.zeros((num_bands, beta_path.shape[1])))
           for pol in pols}
    for k in range(beta_path.shape[1]):
        for pol in pols:
            v, f = bands_at(p, beta_path[:, k], pol, num_bands, sigma)
            out[pol][0][:, k] = v
            out[pol][1][:, k] = f
    return out
