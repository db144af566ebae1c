"""ORACLE (test infrastructure, not product code): CPU restatement of the 2-D scattering hot
path of the reference, /root/reference/Scattering/Scattering_Solver_2D.py.

Restated pieces (each function cites the lines it follows):
  grid / materials      Scattering_Solver_2D.py:28-61
  add_object            :72-86
  plane-wave / point    :89-121
  uniaxial PML          :124-153
  TF/SF mask Q          :156-173
  A_TE / A_TM           :176-188   (explicit scipy sparse, exactly the reference's products)
  matrix-free apply     SURVEY.md §8 a-4 (verified here against the explicit matrices)
  rhs, direct solve     :191-219   (b = (QA - AQ) src ; x = A^-1 b through SuperLU)

The arithmetic below is float64/complex128 numpy + scipy.sparse (sparsetools, SuperLU), the
same third-party code the reference itself executes (scipy, un-pinned in the reference's
README.md:25-27; scipy 1.18.1 in this image).  Parity is pinned by
tests/test_oracle_vs_reference.py (reference imported in-container) and by the golden
fixtures tests/golden/scatter_*.npz generated from the unmodified reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.special

from . import yee_oracle

C0 = 299_792_458.0            # Scattering_Solver_2D.py:24
EPS0 = 8.854187817e-12        # Scattering_Solver_2D.py:25


@dataclass
class ScatterProblem:
    frequency: float
    x_range: float
    y_range: float
    Nx: int
    Ny: int
    omega: float = 0.0
    k0: float = 0.0
    dx: float = 0.0
    dy: float = 0.0
    X: np.ndarray = None
    Y: np.ndarray = None
    mats: dict = field(default_factory=dict)   # ERxx..MRzz, each (Ny, Nx) complex128
    source: np.ndarray = None
    q: np.ndarray = None                       # diagonal of Q, length N (float)

    @property
    def N(self):
        return self.Nx * self.Ny


def make_problem(frequency, x_range, y_range, Nx, Ny) -> ScatterProblem:
    """Scattering_Solver_2D.py:28-61."""
    p = ScatterProblem(float(frequency), float(x_range), float(y_range), int(Nx), int(Ny))
    p.omega = 2 * np.pi * p.frequency
    p.k0 = p.omega / C0
    p.dx, p.dy = p.x_range / p.Nx, p.y_range / p.Ny
    x = (np.arange(p.Nx) + 0.5) * p.dx - p.x_range / 2
    y = (np.arange(p.Ny) + 0.5) * p.dy - p.y_range / 2
    p.X, p.Y = np.meshgrid(x, y, indexing="xy")
    for name in ("ERxx", "ERyy", "ERzz", "MRxx", "MRyy", "MRzz"):
        p.mats[name] = np.ones((p.Ny, p.Nx), dtype=np.complex128)
    p.source = np.zeros(p.N, complex)
    p.q = np.ones(p.N)
    return p


def add_object(p: ScatterProblem, er, mr, region_mask):
    """Scattering_Solver_2D.py:72-86."""
    er = (er,) * 3 if np.isscalar(er) else er
    mr = (mr,) * 3 if np.isscalar(mr) else mr
    for comp, idx in (("xx", 0), ("yy", 1), ("zz", 2)):
        p.mats["ER" + comp][region_mask] = er[idx]
        p.mats["MR" + comp][region_mask] = mr[idx]


def add_source(p: ScatterProblem, src_type="plane_wave", angle_deg=0.0, polarization="TE",
               location=None, amplitude=1.0):
    """Scattering_Solver_2D.py:89-121."""
    kx = p.k0 * np.cos(np.deg2rad(angle_deg))
    ky = p.k0 * np.sin(np.deg2rad(angle_deg))
    if src_type == "plane_wave":
        p.source = (amplitude * np.exp(-1j * (kx * p.X + ky * p.Y))).ravel()
    elif src_type == "point":
        if location is None:
            raise ValueError("For a point source you must supply location=(x0,y0)")
        r = np.hypot(p.X - location[0], p.Y - location[1])
        r[r == 0] = p.dx / 50
        g = scipy.special.hankel1(0, p.k0 * r)
        if polarization.upper() != "TE":
            g = 1j / 4 * g
        p.source = (amplitude * g).ravel()
    else:
        raise ValueError(f"Unknown src_type {src_type}")


def add_upml(p: ScatterProblem, pml_width=20, n=3, sigma_max=5.0, direction="both"):
    """Scattering_Solver_2D.py:124-153."""
    sig_x = np.zeros((p.Ny, p.Nx))
    sig_y = np.zeros((p.Ny, p.Nx))
    for i in range(pml_width):
        s = sigma_max * ((pml_width - i) / pml_width) ** n
        if direction in ("x", "both"):
            sig_x[:, i] = s
            sig_x[:, -i - 1] = s
        if direction in ("y", "both"):
            sig_y[i, :] = s
            sig_y[-i - 1, :] = s
    Sx = 1.0 + 1j * sig_x / (EPS0 * p.omega)
    Sy = 1.0 + 1j * sig_y / (EPS0 * p.omega)
    for pre in ("ER", "MR"):
        p.mats[pre + "xx"] *= Sy / Sx
        p.mats[pre + "yy"] *= Sx / Sy
        p.mats[pre + "zz"] *= Sx * Sy


def add_mask(p: ScatterProblem, value=30):
    """Scattering_Solver_2D.py:156-173 (scalar and array forms)."""
    if np.isscalar(value):
        v = int(value)
        q = np.ones((p.Ny, p.Nx))
        if 2 * v < min(p.Nx, p.Ny):
            q[v:-v, v:-v] = 0
        p.q = q.ravel()
    else:
        arr = np.asarray(value)
        if arr.shape != (p.Ny, p.Nx):
            raise ValueError(f"mask must be shape {(p.Ny, p.Nx)}")
        p.q = arr.ravel()


def operators(p: ScatterProblem):
    """Scattering_Solver_2D.py:63-69: normalised resolution k0*dx, Dirichlet on both axes."""
    return yee_oracle.yeeder2d([p.Nx, p.Ny], [p.k0 * p.dx, p.k0 * p.dy], [0, 0])


def build_A(p: ScatterProblem, pol: str):
    """Explicit sparse system matrix, Scattering_Solver_2D.py:176-181 (TE) / :183-188 (TM)."""
    DEX, DEY, DHX, DHY = operators(p)
    if pol.upper() == "TE":
        inv_a = sp.diags(p.mats["MRyy"].ravel()).power(-1)
        inv_b = sp.diags(p.mats["MRxx"].ravel()).power(-1)
        dg = sp.diags(p.mats["ERzz"].ravel())
        return DHX @ inv_a @ DEX + DHY @ inv_b @ DEY + dg
    inv_a = sp.diags(p.mats["ERyy"].ravel()).power(-1)
    inv_b = sp.diags(p.mats["ERxx"].ravel()).power(-1)
    dg = sp.diags(p.mats["MRzz"].ravel())
    return DEX @ inv_a @ DHX + DEY @ inv_b @ DHY + dg


def stencil_coefficients(p: ScatterProblem, pol: str):
    """(ca, cb, cd, inv_dx2, inv_dy2) of the matrix-free form in SURVEY.md §8 a-4:
    ca = 1/m_a, cb = 1/m_b with (m_a, m_b, diag) = (MRyy, MRxx, ERzz) for TE and
    (ERyy, ERxx, MRzz) for TM; 1/dx^2 with dx = k0*dx (Scattering_Solver_2D.py:65-66)."""
    if pol.upper() == "TE":
        ma, mb, dg = p.mats["MRyy"], p.mats["MRxx"], p.mats["ERzz"]
    else:
        ma, mb, dg = p.mats["ERyy"], p.mats["ERxx"], p.mats["MRzz"]
    dxn, dyn = p.k0 * p.dx, p.k0 * p.dy
    return 1.0 / ma, 1.0 / mb, dg.copy(), 1.0 / (dxn * dxn), 1.0 / (dyn * dyn)


def apply_stencil(pol, ca, cb, cd, inv_dx2, inv_dy2, x):
    """y = A x without forming A (Dirichlet walls).  x, y: (Ny, Nx) complex.

    TE:  DHX ca DEX + DHY cb DEY + cd : forward difference first (value beyond the far wall is
         0), coefficient at the *same* cell, then backward difference (value before the near
         wall is 0).
    TM:  DEX ca DHX + DEY cb DHY + cd : backward difference first, coefficient, then forward.
    """
    x = np.asarray(x).reshape(ca.shape)
    Ny, Nx = x.shape
    y = cd * x
    if pol.upper() == "TE":
        fx = np.zeros_like(x)
        fx[:, :-1] = x[:, 1:]
        fx = ca * (fx - x)                      # ca_ij (e_{i+1,j} - e_ij), e_Nx = 0
        y += inv_dx2 * fx
        y[:, 1:] -= inv_dx2 * fx[:, :-1]
        fy = np.zeros_like(x)
        fy[:-1, :] = x[1:, :]
        fy = cb * (fy - x)
        y += inv_dy2 * fy
        y[1:, :] -= inv_dy2 * fy[:-1, :]
    else:
        gx = x.copy()
        gx[:, 1:] -= x[:, :-1]
        gx = ca * gx                            # ca_ij (h_ij - h_{i-1,j}), h_-1 = 0
        y -= inv_dx2 * gx
        y[:, :-1] += inv_dx2 * gx[:, 1:]        # forward term vanishes at the far wall
        gy = x.copy()
        gy[1:, :] -= x[:-1, :]
        gy = cb * gy
        y -= inv_dy2 * gy
        y[:-1, :] += inv_dy2 * gy[1:, :]
    return y


def rhs(A, p: ScatterProblem):
    """b = (Q A - A Q) src, Scattering_Solver_2D.py:198 / :213."""
    Q = sp.diags(p.q, format="csr")
    return np.asarray((Q @ A - A @ Q) @ p.source[:, None]).ravel()


def rhs_matrix_free(p: ScatterProblem, pol: str):
    """Same right-hand side through the stencil: Q(A s) - A(Q s)."""
    co = stencil_coefficients(p, pol)
    shape = (p.Ny, p.Nx)
    q = p.q.reshape(shape)
    s = p.source.reshape(shape)
    return (q * apply_stencil(pol, *co, s) - apply_stencil(pol, *co, q * s)).ravel()


def solve_total_field(p: ScatterProblem, pol: str, A=None):
    """Scattering_Solver_2D.py:191-219: SuperLU direct solve, result reshaped (Ny, Nx)."""
    if A is None:
        A = build_A(p, pol)
    b = rhs(A, p)
    x = spla.spsolve(A.tocsc(), b)
    return x.reshape(p.Ny, p.Nx)


def cylinder_problem(N, eps_r=-1e6, f0=10e9, pml_width=40, mask=80, sigma_max=10, angle=45.0):
    """The reference example scaled to N x N at 50 cells per wavelength
    (Scattering/example_scattering_by_cylinder.py:8-32; SURVEY.md §8d synthetic inputs)."""
    lam = 299792458 / f0
    L = (N / 50.0) * lam
    p = make_problem(f0, L, L, N, N)
    add_object(p, eps_r, 1.0, (p.X ** 2 + p.Y ** 2) <= (0.5 * lam) ** 2)
    add_upml(p, pml_width=pml_width, sigma_max=sigma_max)
    add_mask(p, mask)
    add_source(p, "plane_wave", angle_deg=angle, polarization="TE")
    return p
