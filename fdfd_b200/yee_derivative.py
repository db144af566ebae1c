"""Drop-in for the reference's ``yee_derivative.yeeder2d`` with the CSR assembly done by the
sm_100a kernel K1 (csrc/yee_assemble.cu) behind the C-ABI ``fdfd_yee2d_csr_fill*``.

Reference: Scattering/yee_derivative.py:6-74 (identical copy in Band_Diagram_Solver/).
Same signature, same return contract: ``(DEX, DEY, DHX, DHY)`` with DEX/DEY ``csr_matrix``,
DHX/DHY ``csc_matrix`` (= -D^H), int32 index arrays, float64 data on a Dirichlet axis and
complex128 on a Bloch (or single-cell) axis — bit-identical to what scipy produces.

The <=3 distinct values per axis are evaluated here with the reference's own numpy
expressions (true division by dx; ``np.exp(-1j*k*N*d)/d``) and handed to the kernel, which
only scatters them: that is what makes the *values* bit-exact too.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib


def _axis_values(N, d, bc, k):
    """(is_complex, [diag_re, diag_im, fwd_re, fwd_im, wrap_re, wrap_im]) for one axis."""
    if N == 1:
        v = (np.ones(1) * (-1j * k))[0]              # yee_derivative.py:47-48 / :60-61
        return True, [v.real, v.imag, 0.0, 0.0, 0.0, 0.0]
    base = np.array([-1.0, 1.0]) / d                 # yee_derivative.py:53 / :65
    if bc == 1:
        w = np.exp(-1j * k * N * d) / d              # yee_derivative.py:56 / :67
        return True, [base[0], 0.0, base[1], 0.0, w.real, w.imag]
    return False, [base[0], 0.0, base[1], 0.0, 0.0, 0.0]


def _check_args(NS, RES, BC):
    Nx, Ny = int(NS[0]), int(NS[1])
    if Nx < 1 or Ny < 1:
        raise ValueError("NS must hold positive grid sizes")
    bc = [int(BC[0]), int(BC[1])]
    if any(b not in (0, 1) for b in bc):
        raise ValueError("BC entries must be 0 (Dirichlet) or 1 (periodic)")
    return Nx, Ny, RES[0], RES[1], bc


def yeeder2d(NS, RES, BC, kinc=None):
    """Derivative matrices on a 2-D Yee grid (GPU assembly, scipy return types)."""
    _lib.require_cuda()
    L = _lib.lib()
    Nx, Ny, dx, dy, bc = _check_args(NS, RES, BC)
    if kinc is None:
        kinc = np.array([0, 0])
    M = Nx * Ny
    nx, ny = C.c_int64(), C.c_int64()
    cx, cy = C.c_int(), C.c_int()
    _lib.check(L.fdfd_yee2d_csr_nnz(Nx, Ny, bc[0], bc[1], C.byref(nx), C.byref(ny), C.byref(cx), C.byref(cy)))
    out = []
    for axis, (N, d, nnz) in enumerate(((Nx, dx, nx.value), (Ny, dy, ny.value))):
        is_c, vals = _axis_values(N, d, bc[axis], kinc[axis])
        dtype = np.complex128 if is_c else np.float64
        indptr = np.empty(M + 1, dtype=np.int32)
        indices = np.empty(nnz, dtype=np.int32)
        e_data = np.empty(nnz, dtype=dtype)
        h_data = np.empty(nnz, dtype=dtype)
        cvals = (C.c_double * 6)(*[float(v) for v in vals])
        _lib.check(L.fdfd_yee2d_csr_fill_host(Nx, Ny, axis, bc[axis], cvals, indptr.ctypes.data,
                                              indices.ctypes.data, e_data.ctypes.data, h_data.ctypes.data))
        out.append((indptr, indices, e_data, h_data))
    (px, ix, ex, hx), (py, iy, ey, hy) = out
    DEX = sp.csr_matrix((ex, ix, px), shape=(M, M))
    DEY = sp.csr_matrix((ey, iy, py), shape=(M, M))
    DHX = sp.csc_matrix((hx, ix.copy(), px.copy()), shape=(M, M))
    DHY = sp.csc_matrix((hy, iy.copy(), py.copy()), shape=(M, M))
    return DEX, DEY, DHX, DHY


def yeeder2d_device(NS, RES, BC, kinc=None, device=None):
    """Same operators left on the GPU: a dict of torch tensors per axis
    ``{'indptr','indices','e_data','h_data'}`` (CSR arrays of DEX/DEY; DHX/DHY are the CSC
    reading of the same index arrays with ``h_data``)."""
    import torch

    _lib.require_cuda()
    L = _lib.lib()
    Nx, Ny, dx, dy, bc = _check_args(NS, RES, BC)
    if kinc is None:
        kinc = np.array([0, 0])
    device = torch.device(device or "cuda")
    M = Nx * Ny
    nx, ny = C.c_int64(), C.c_int64()
    _lib.check(L.fdfd_yee2d_csr_nnz(Nx, Ny, bc[0], bc[1], C.byref(nx), C.byref(ny), None, None))
    stream = torch.cuda.current_stream(device).cuda_stream
    res = {}
    with torch.cuda.device(device):
        for axis, name, N, d, nnz in ((0, "x", Nx, dx, nx.value), (1, "y", Ny, dy, ny.value)):
            is_c, vals = _axis_values(N, d, bc[axis], kinc[axis])
            dt = torch.complex128 if is_c else torch.float64
            t = dict(indptr=torch.empty(M + 1, dtype=torch.int32, device=device),
                     indices=torch.empty(nnz, dtype=torch.int32, device=device),
                     e_data=torch.empty(nnz, dtype=dt, device=device),
                     h_data=torch.empty(nnz, dtype=dt, device=device))
            cvals = (C.c_double * 6)(*[float(v) for v in vals])
            _lib.check(L.fdfd_yee2d_csr_fill(Nx, Ny, axis, bc[axis], cvals, t["indptr"].data_ptr(),
                                             t["indices"].data_ptr(), t["e_data"].data_ptr(),
                                             t["h_data"].data_ptr(), stream))
            res[name] = t
    return res
