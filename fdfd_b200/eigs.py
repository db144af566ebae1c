"""K6 — GPU shift-invert Krylov-Schur (restarted Arnoldi) eigen-solver.

Replaces the ARPACK calls of the reference — ``eigs(A, M=M, k, sigma)`` at
Band_Diagram_Solver.py:313,320 (ARPACK mode 3: Arnoldi on ``(A - sigma M)^-1 M`` with a SuperLU
factorisation) — by:
  * the Krylov basis kept on the device, orthogonalised with the fused tall-skinny kernels of K4
    (``fdfd_blas_multi_dot`` / ``fdfd_blas_multi_axpy``: classical Gram-Schmidt with a second
    pass, one read of the basis per pass, in-kernel reductions);
  * the shifted solves done on the GPU (iterative through K3 on the matrix-free operator, or a
    dense LU for unit cells of a few thousand unknowns);
  * Stewart's Krylov-Schur restart; only the (ncv x ncv) projected matrix visits the host.
Eigenvalues of the pencil are recovered as ``lambda = sigma + 1/mu``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.linalg as sla

from . import _lib


class DeviceBlas:
    """Handle of the K4 workspace used by the eigen-solver."""

    def __init__(self, device=None):
        import torch
        _lib.require_cuda()
        self._L = _lib.lib()
        self.device = torch.device(device or "cuda")
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_blas_create(C.byref(self._h)))

    def close(self):
        if self._h and self._h.value:
            self._L.fdfd_blas_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _code(t):
        import torch
        return _lib.C128 if t.dtype == torch.complex128 else _lib.C64

    def _stream(self):
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream

    def multi_dot(self, V, k, w):
        """h = V[:k]^H w  (V: (m, n) device tensor with contiguous rows) -> complex numpy (k,)."""
        import torch
        h = np.empty(k, dtype=np.complex128)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_blas_multi_dot(self._h, self._code(w), w.numel(), k, V.data_ptr(), V.stride(0),
                                                   w.data_ptr(), h.ctypes.data, self._stream()))
        return h

    def multi_axpy(self, V, k, w, h, sign=-1.0, want_norm=False):
        """w += sign * sum_i h[i] V[i]; returns ||w||^2 when asked."""
        import torch
        h = np.ascontiguousarray(h, dtype=np.complex128)
        nrm = C.c_double(0.0)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_blas_multi_axpy(self._h, self._code(w), w.numel(), k, V.data_ptr(), V.stride(0),
                                                    w.data_ptr(), h.ctypes.data, float(sign),
                                                    C.byref(nrm) if want_norm else None, self._stream()))
        return nrm.value if want_norm else None


def krylov_schur(op, n, k, *, ncv=None, tol=1e-10, max_restarts=100, v0=None, seed=0, device=None,
                 dtype=None, blas=None):
    """Largest-magnitude eigenpairs of the linear operator ``op`` (device tensor -> device tensor).

    Returns (mu, X, resid, restarts): eigenvalues (k,), Ritz vectors as a (k, n) device tensor
    with unit 2-norm rows, the Arnoldi residual estimates |b^T s_i| and the number of restarts.
    """
    import torch

    device = torch.device(device or "cuda")
    dtype = dtype or torch.complex128
    if ncv is None:
        ncv = max(2 * k + 1, 20)              # ARPACK's default (scipy arpack.py:331-336)
    ncv = int(min(max(ncv, k + 2), n))
    blas = blas or DeviceBlas(device)
    V = torch.zeros((ncv + 1, n), dtype=dtype, device=device)
    H = np.zeros((ncv + 1, ncv), dtype=np.complex128)
    if v0 is None:
        rng = np.random.default_rng(seed)
        v0 = rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)
        v0 = torch.from_numpy(v0).to(device=device, dtype=dtype)
    V[0] = v0 / torch.linalg.vector_norm(v0)
    p = 0
    keep = min(ncv - 1, k + max(1, (ncv - k) // 2))
    restarts = 0
    while True:
        for j in range(p, ncv):
            w = op(V[j]).reshape(-1).contiguous()
            h = blas.multi_dot(V, j + 1, w)
            blas.multi_axpy(V, j + 1, w, h, -1.0)
            h2 = blas.multi_dot(V, j + 1, w)                      # second Gram-Schmidt pass
            nrm2 = blas.multi_axpy(V, j + 1, w, h2, -1.0, want_norm=True)
            H[:j + 1, j] = h + h2
            beta = float(np.sqrt(max(nrm2, 0.0)))
            H[j + 1, j] = beta
            if beta <= 1e-300:
                # invariant subspace: continue with a fresh random direction
                w = torch.randn(n, dtype=torch.float64, device=device).to(dtype)
                h = blas.multi_dot(V, j + 1, w)
                nrm2 = blas.multi_axpy(V, j + 1, w, h, -1.0, want_norm=True)
                beta_r = float(np.sqrt(nrm2))
                V[j + 1] = w / beta_r
            else:
                V[j + 1] = w / beta
        Hm = H[:ncv, :ncv]
        b = H[ncv, :ncv].copy()
        mu, S = np.linalg.eig(Hm)
        order = np.argsort(-np.abs(mu))
        mu, S = mu[order], S[:, order]
        S = S / np.linalg.norm(S, axis=0, keepdims=True)
        res = np.abs(b @ S)
        converged = np.all(res[:k] <= tol * np.abs(mu[:k]))
        if converged or restarts >= max_restarts:
            Sk = torch.from_numpy(np.ascontiguousarray(S[:, :k].T)).to(device=device, dtype=dtype)
            X = Sk @ V[:ncv]
            X = X / torch.linalg.vector_norm(X, dim=1, keepdim=True)
            return mu[:k], X, res[:k], restarts
        # Krylov-Schur restart: keep the `keep` dominant Schur vectors
        thr = np.sort(np.abs(mu))[::-1][keep - 1]
        T, Q, sdim = sla.schur(Hm, output="complex", sort=lambda x: abs(x) >= thr * (1 - 1e-12))
        pk = min(max(sdim, k), ncv - 1)
        Qk = torch.from_numpy(np.ascontiguousarray(Q[:, :pk].T)).to(device=device, dtype=dtype)
        Vnew = Qk @ V[:ncv]
        last = V[ncv].clone()
        V[:pk] = Vnew
        V[pk] = last
        H[:] = 0
        H[:pk, :pk] = T[:pk, :pk]
        H[pk, :pk] = b @ Q[:, :pk]
        p = pk
        restarts += 1


def eigs_shift_invert(apply_A, apply_M, solve_shifted, n, k, sigma, *, ncv=None, tol=1e-10, max_restarts=100,
                      v0=None, seed=0, device=None, dtype=None, blas=None):
    """k eigenvalues of ``A x = lambda M x`` closest to ``sigma``.

    apply_A / apply_M: device tensor -> device tensor; solve_shifted(rhs) returns
    ``(A - sigma M)^-1 rhs``.  Returns (lambda (k,), X (k, n) device, true relative residuals).
    """
    import torch

    def op(v):
        return solve_shifted(apply_M(v))

    mu, X, _, restarts = krylov_schur(op, n, k, ncv=ncv, tol=tol, max_restarts=max_restarts, v0=v0, seed=seed,
                                      device=device, dtype=dtype, blas=blas)
    lam = sigma + 1.0 / mu
    resid = np.zeros(k)
    for i in range(k):
        ax = apply_A(X[i]).reshape(-1)
        mx = apply_M(X[i]).reshape(-1)
        lam_t = torch.tensor(lam[i], dtype=X.dtype, device=X.device)
        num = torch.linalg.vector_norm(ax - lam_t * mx)
        den = torch.linalg.vector_norm(ax) + abs(lam[i]) * torch.linalg.vector_norm(mx)
        resid[i] = float(num / den) if float(den) > 0 else float(num)
    return lam, X, resid, restarts
