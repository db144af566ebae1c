"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's
``yeeder2d`` — /root/reference/Scattering/yee_derivative.py:6-74 (byte-identical function at
Band_Diagram_Solver/yee_derivative.py:6-74).

The reference goes through scipy's DIA->CSR conversion and CSR addition; this restatement
writes the CSR arrays in closed form with numpy so that the *structure* of the result is
spelled out (and can be checked against scipy's bytes):

* flat index n = i + Nx*j, x fastest (Scattering_Solver_2D.py:41,61,107).
* Dirichlet x (BC[0]==0, Nx>1): row n holds (n, -1/dx) and, unless i==Nx-1, (n+1, +1/dx)
  (yee_derivative.py:50-53; the explicit zero written at :52 is dropped by dia->csr).
* Bloch x (BC[0]==1, Nx>1): rows with i==Nx-1 additionally hold
  (n-(Nx-1), exp(-1j*kx*Nx*dx)/dx) (yee_derivative.py:54-57); columns sorted, so the wrap
  entry comes first in that row.  dtype becomes complex128.
* Nx==1: DEX = (-1j*kx) * I, complex128 (yee_derivative.py:47-48).
* y axis: same with stride Nx and wrap column n-(M-Nx) for j==Ny-1 (yee_derivative.py:59-68).
* DHX = -DEX^H, DHY = -DEY^H (yee_derivative.py:71-72): CSC whose (indptr, indices) are the
  CSR arrays of DEX/DEY and whose data is -conj(data).

Values are evaluated with the very same numpy scalar expressions the reference uses (true
division ``-1.0/dx``; ``np.exp(-1j*k*N*d)/d``) so that data bytes match, not just indices.

Parity pinned by tests/test_oracle_vs_reference.py (reference imported in-container) and
by the fixtures in tests/golden/yee2d_*.npz (SURVEY.md §8c known-answer values).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def axis_scalars(N, d, bc, k):
    """The <=3 distinct values one axis of yeeder2d can contain, evaluated exactly as the
    reference does (yee_derivative.py:47-57 / :60-68).  Returns (is_complex, diag, fwd, wrap)."""
    if N == 1:
        v = (np.ones(1) * (-1j * k))[0]          # eye(M).data * scalar  (scipy _mul_scalar)
        return True, complex(v), 0.0, 0.0
    base = np.array([-1.0, 1.0]) / d             # diags([d0,d1]).data / dx
    if bc == 1:
        wrap = np.exp(-1j * k * N * d) / d
        # float CSR + complex CSR -> complex128; real entries become (v + 0j)
        return True, complex(base[0]), complex(base[1]), complex(wrap)
    return False, float(base[0]), float(base[1]), 0.0


def _axis_csr(Nx, Ny, axis, d, bc, k):
    """CSR arrays (indptr int32, indices int32, data) of DEX (axis=0) or DEY (axis=1)."""
    M = Nx * Ny
    N = Nx if axis == 0 else Ny
    is_c, diag, fwd, wrap = axis_scalars(N, d, bc, k)
    dtype = np.complex128 if is_c else np.float64
    n = np.arange(M, dtype=np.int64)
    if N == 1:
        indptr = np.arange(M + 1, dtype=np.int32)
        return indptr, n.astype(np.int32), np.full(M, diag, dtype=dtype)
    if axis == 0:
        pos = n % Nx            # i
        stride, span = 1, Nx - 1
    else:
        pos = n // Nx           # j
        stride, span = Nx, M - Nx
    last = pos == N - 1
    periodic = bc == 1
    cnt = np.where(last, 2 if periodic else 1, 2)
    indptr64 = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(cnt, out=indptr64[1:])
    nnz = int(indptr64[-1])
    indices = np.empty(nnz, dtype=np.int64)
    data = np.empty(nnz, dtype=dtype)
    s = indptr64[:-1]
    inner = ~last
    # interior rows: [n, n+stride]
    indices[s[inner]] = n[inner]
    data[s[inner]] = diag
    indices[s[inner] + 1] = n[inner] + stride
    data[s[inner] + 1] = fwd
    if periodic:
        # wrap rows: [n-span, n]  (wrap column is smaller -> sorted first)
        indices[s[last]] = n[last] - span
        data[s[last]] = wrap
        indices[s[last] + 1] = n[last]
        data[s[last] + 1] = diag
    else:
        indices[s[last]] = n[last]
        data[s[last]] = diag
    return indptr64.astype(np.int32), indices.astype(np.int32), data


def yeeder2d_arrays(NS, RES, BC, kinc=None):
    """Closed-form CSR arrays for DEX and DEY: ((indptr, indices, data), (indptr, indices, data))."""
    Nx, Ny = int(NS[0]), int(NS[1])
    dx, dy = RES
    if kinc is None:
        kinc = np.array([0, 0])
    ex = _axis_csr(Nx, Ny, 0, dx, BC[0], kinc[0])
    ey = _axis_csr(Nx, Ny, 1, dy, BC[1], kinc[1])
    return ex, ey


def yeeder2d(NS, RES, BC, kinc=None):
    """Same return contract as the reference: (DEX csr, DEY csr, DHX csc, DHY csc)."""
    Nx, Ny = int(NS[0]), int(NS[1])
    M = Nx * Ny
    (px, ix, vx), (py, iy, vy) = yeeder2d_arrays(NS, RES, BC, kinc)
    DEX = sp.csr_matrix((vx, ix, px), shape=(M, M))
    DEY = sp.csr_matrix((vy, iy, py), shape=(M, M))
    DHX = sp.csc_matrix((-np.conj(vx), ix.copy(), px.copy()), shape=(M, M))
    DHY = sp.csc_matrix((-np.conj(vy), iy.copy(), py.copy()), shape=(M, M))
    return DEX, DEY, DHX, DHY
