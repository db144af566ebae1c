"""GPU drop-in for the reference's ``BandDiagramSolver2D``
(Band_Diagram_Solver/Band_Diagram_Solver.py:58-560): same constructor, geometry helpers, Bloch
path utilities and ``compute_band_structure`` contract (``BandStructureResult`` with
``frequencies[pol]`` (num_bands, N) = a/lambda and ``eigenvalues[pol]`` = (omega/c)^2).

What changed underneath:
  * per Bloch vector the reference rebuilds four sparse operators with ``yeeder2d`` and two
    sparse products (Band_Diagram_Solver.py:304-319); here the pencil is the matrix-free K2
    operator (only the two Bloch phases change between path points);
  * ``eigs(A, M=M, k, sigma)`` (ARPACK mode 3 + SuperLU, :313/:320) is replaced by the GPU
    one-CTA-per-path-point kernels of ``csrc/band.cu`` (block LU of the cyclic block-tridiagonal
    shifted operator in shared memory + Arnoldi / Krylov-Schur cycles, all path points in lock-step);
    cells those kernels do not cover fall back to the Krylov-Schur solver of ``fdfd_b200/eigs.py``
    around a dense batched inverse or the K3 Krylov solver.
For the reference's default ``eig_sigma = 0`` the shift actually used is ``-delta`` (delta > 0): the
pencil's eigenvalues do not depend on it, and it keeps the shifted operator non-singular at Gamma where
``sigma = 0`` is exactly singular (the reference relies on SuperLU surviving that).  A non-zero
``eig_sigma`` is used as given.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Iterable, Sequence

import numpy as np

from . import _lib
from . import eigs as _eigs
from .operators import HelmholtzOperator2D

MaskLike = "np.ndarray | Callable[[np.ndarray, np.ndarray], np.ndarray]"


@dataclass
class BandStructureResult:
    """Same container as the reference (Band_Diagram_Solver.py:30-55)."""
    beta_path: np.ndarray
    tick_positions: list
    tick_labels: list
    frequencies: dict
    eigenvalues: dict


class BandDiagramSolver2D:
    #: unit cells up to this many unknowns use the dense-LU shifted solve
    DENSE_LIMIT = 6000
    #: dense assembly of the unit-cell operator: "columns" = one stencil apply per unit vector,
    #: "probe" = coloured probing (25 applies for a 40x40 cell; validated on the host against explicit
    #: matrices in tests/test_eigs_host.py, to become the default once timed on hardware)
    ASSEMBLY = "columns"

    def __init__(self, a, Nx, Ny=None, *, b=None, background_er=1.0, background_ur=1.0,
                 boundary_conditions=(1, 1), device=None):
        self.a = float(a)
        self.b = float(b) if b is not None else float(a)
        self.Nx = int(Nx)
        self.Ny = int(Ny) if Ny is not None else int(Nx)
        self.boundary_conditions = tuple(boundary_conditions)
        self.dx, self.dy = self.a / self.Nx, self.b / self.Ny
        self.Nx2, self.Ny2 = 2 * self.Nx, 2 * self.Ny
        self.dx2, self.dy2 = self.dx / 2, self.dy / 2
        xa2 = np.arange(1, self.Nx2 + 1) * self.dx2
        ya2 = np.arange(1, self.Ny2 + 1) * self.dy2
        self.xa2 = xa2 - np.mean(xa2)
        self.ya2 = ya2 - np.mean(ya2)
        self.X2, self.Y2 = np.meshgrid(self.xa2, self.ya2, indexing="ij")
        self.ER2 = np.full((self.Nx2, self.Ny2), background_er, dtype=complex)
        self.UR2 = np.full((self.Nx2, self.Ny2), background_ur, dtype=complex)
        self._beta_path = None
        self._results = None
        self.device = device
        self.solver_info = {}

    # ---- geometry (host, as the reference: Band_Diagram_Solver.py:119-181) ---------------------
    def add_object(self, mask, *, er=None, ur=None):
        sel = np.asarray(mask(self.X2, self.Y2) if callable(mask) else mask, dtype=bool)
        if sel.shape != self.ER2.shape:
            raise ValueError("Mask must match the helper grid shape (2× resolution).")
        for name, val in (("ER2", er), ("UR2", ur)):
            if val is None:
                continue
            arr = np.asarray(val, dtype=complex)
            if arr.shape not in ((), sel.shape):
                raise ValueError(f"{name[:2].lower()} must be a scalar or have the same shape as the mask.")
            setattr(self, name, np.where(sel, arr, getattr(self, name)))

    def add_circular_inclusion(self, radius, *, center=(0.0, 0.0), er=None, ur=None):
        cx, cy = center
        self.add_object((self.X2 - cx) ** 2 + (self.Y2 - cy) ** 2 <= radius ** 2, er=er, ur=ur)

    # ---- Bloch path (Band_Diagram_Solver.py:183-251, :338-355) --------------------------------------
    def default_rectangular_lattice_path(self):
        gx, gy = 2 * np.pi / self.a, 2 * np.pi / self.b
        pts = [np.array([0.0, 0.0]), 0.5 * np.array([gx, 0.0]), 0.5 * np.array([gx, gy]), np.array([0.0, 0.0])]
        return pts, ["Γ", "X", "M", "Γ"]

    def generate_bloch_path(self, symmetry_points, total_points):
        if total_points < len(symmetry_points):
            raise ValueError("total_points must be at least the number of symmetry points.")
        pts = np.asarray(symmetry_points, dtype=float)
        if pts.ndim != 2 or pts.shape[1] != 2:
            raise ValueError("symmetry_points must be an iterable of 2D coordinates.")
        seg = np.linalg.norm(np.diff(pts, axis=0), axis=1)
        total = seg.sum()
        if total == 0:
            seg = np.ones_like(seg)
            total = seg.sum()
        remaining = total_points - 1
        counts = [max(1, int(round(remaining * (length / total)))) for length in seg]
        diff, idx = remaining - sum(counts), 0
        while diff != 0 and counts:
            c = idx % len(counts)
            if diff > 0:
                counts[c] += 1
                diff -= 1
            elif counts[c] > 1:
                counts[c] -= 1
                diff += 1
            idx += 1
        betas, ticks, acc = [pts[0]], [0], 0
        for start, stop, n_seg in zip(pts[:-1], pts[1:], counts):
            betas.extend(start + (s / n_seg) * (stop - start) for s in range(1, n_seg + 1))
            acc += n_seg
            ticks.append(acc)
        beta_path = np.column_stack(betas)
        self._beta_path = beta_path
        self._tick_positions = list(ticks)
        return beta_path, ticks

    def set_tick_labels(self, labels, positions):
        if len(labels) != len(positions):
            raise ValueError("labels and positions must have the same length.")
        self._tick_labels = list(labels)
        self._tick_positions = list(positions)

    def _tick_positions_from_path(self, beta_path):
        if self._beta_path is None or not np.array_equal(beta_path, self._beta_path):
            return list(range(beta_path.shape[1]))
        return list(getattr(self, "_tick_positions", range(beta_path.shape[1])))

    # ---- materials on the Yee grid (Band_Diagram_Solver.py:536-550) ---------------------------------
    def _yee_tensors(self):
        return dict(ERxx=self.ER2[1::2, ::2], ERyy=self.ER2[::2, 1::2], ERzz=self.ER2[::2, 1::2],
                    URxx=self.UR2[1::2, ::2], URyy=self.UR2[::2, 1::2], URzz=self.UR2[::2, 1::2])

    def _pencil_arrays(self, pol):
        """(K2 operator type, ca, cb, M) as (Ny, Nx) arrays: the reference flattens its (Nx, Ny)
        arrays with order='F' (:286-291), i.e. n = i + Nx*j = C-order of the transpose."""
        t = self._yee_tensors()
        tr = lambda a: np.ascontiguousarray(a.T)  # noqa: E731
        if pol == "TM":   # A_tm = -DHX URyy^-1 DEX - DHY URxx^-1 DEY, M = ERzz   (:312-313)
            return "TE", tr(1.0 / t["URyy"]), tr(1.0 / t["URxx"]), tr(t["ERzz"])
        return "TM", tr(1.0 / t["ERyy"]), tr(1.0 / t["ERxx"]), tr(t["URzz"])  # (:319-320)

    def _operator(self, kind, ca, cb, cd, phase):
        return HelmholtzOperator2D(kind, self.Nx, self.Ny, ca, cb, cd, -1.0 / self.dx ** 2, -1.0 / self.dy ** 2,
                                   bc=self.boundary_conditions, phase=phase, device=self.device)

    def _phases(self, beta):
        # exp(-1j*k*N*d) exactly as yeeder2d evaluates it (yee_derivative.py:56,67)
        return (np.exp(-1j * beta[0] * self.Nx * self.dx), np.exp(-1j * beta[1] * self.Ny * self.dy))

    def _dense(self, op):
        """Dense matrix of a K2 operator: one stencil apply per unit vector (unit cells only)."""
        import torch
        n = op.n
        D = torch.empty((n, n), dtype=op.tdtype, device=op.device)
        e = torch.zeros(n, dtype=op.tdtype, device=op.device)
        for c in range(n):
            e[c] = 1.0
            op.apply(e, out=D[c])          # row c of D^T = column c of the operator
            e[c] = 0.0
        return D.T.contiguous()

    @staticmethod
    def _probe_period(N, periodic):
        """Smallest colour period >= 3 along an axis of N points such that points with equal colour are
        at least 3 apart, also across the Bloch wrap (which needs the period to divide N)."""
        for p in range(3, N + 1):
            if not periodic or N % p == 0:
                return p
        return None

    def _dense_probed(self, op):
        """Dense matrix of a five-point operator from a few stencil applies (sparse-Jacobian probing):
        columns whose stencils cannot overlap share one probe vector, colour (i mod px, j mod py) with
        px, py >= 3 (dividing Nx, Ny on Bloch axes), so px*py applies replace Nx*Ny; every non-zero of
        column c is read back from the probe of c's colour at the rows of c's stencil.  Falls back to one
        apply per column when an axis is too short to colour."""
        import torch
        Nx, Ny, n = op.Nx, op.Ny, op.n
        bcx, bcy = (int(b) for b in op.bc)
        px, py = self._probe_period(Nx, bcx == 1), self._probe_period(Ny, bcy == 1)
        if px is None or py is None or n != Nx * Ny:
            return self._dense(op)
        dev, tdt = op.device, op.tdtype
        idx = torch.arange(n, device=dev)
        ci, cj = idx % Nx, idx // Nx
        colour = (ci % px) + px * (cj % py)
        Y = torch.empty((px * py, n), dtype=tdt, device=dev)
        for c in range(px * py):
            op.apply((colour == c).to(tdt), out=Y[c])
        D = torch.zeros((n, n), dtype=tdt, device=dev)
        for di, dj in ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)):
            ri, rj = ci + di, cj + dj
            ok = torch.ones(n, dtype=torch.bool, device=dev)
            if bcx == 1:
                ri = ri % Nx
            else:
                ok &= (ri >= 0) & (ri < Nx)
            if bcy == 1:
                rj = rj % Ny
            else:
                ok &= (rj >= 0) & (rj < Ny)
            r = (ri + Nx * rj)[ok]
            c = idx[ok]
            D[r, c] = Y[colour[ok], r]
        return D

    def _dense_pieces(self, kind, ca, cb, cdev):
        """The shifted operator is affine in (phx, conj phx, phy, conj phy):
        S(phx, phy) = C0 + phx*Cx + conj(phx)*Cx' + phy*Cy + conj(phy)*Cy'.  Five assemblies at fixed
        phases determine the pieces once per polarisation; every path point is then a 5-term
        elementwise combination on the device."""
        dense = self._dense_probed if self.ASSEMBLY == "probe" else self._dense
        S = {ph: dense(self._operator(kind, ca, cb, cdev, ph))
             for ph in ((1, 1), (-1, 1), (1j, 1), (1, -1), (1, 1j))}
        sx_sum = (S[(1, 1)] - S[(-1, 1)]) / 2            # Cx + Cx'
        base_x = (S[(1, 1)] + S[(-1, 1)]) / 2            # C0 + Cy + Cy'
        sx_dif = (S[(1j, 1)] - base_x) / 1j              # Cx - Cx'
        sy_sum = (S[(1, 1)] - S[(1, -1)]) / 2
        base_y = (S[(1, 1)] + S[(1, -1)]) / 2            # C0 + Cx + Cx'
        sy_dif = (S[(1, 1j)] - base_y) / 1j
        Cx, Cxp = (sx_sum + sx_dif) / 2, (sx_sum - sx_dif) / 2
        Cy, Cyp = (sy_sum + sy_dif) / 2, (sy_sum - sy_dif) / 2
        C0 = S[(1, 1)] - sx_sum - sy_sum
        return C0, Cx, Cxp, Cy, Cyp

    # ---- solver core (Band_Diagram_Solver.py:256-336) ------------------------------------------------
    def compute_band_structure(self, beta_path, *, num_bands, polarisations: Iterable[str] = ("TE", "TM"),
                               eig_sigma=0.0, tol=1e-11, inner="auto", batch=None):
        import torch

        polarisations = tuple(p.upper() for p in polarisations)
        allowed = {"TE", "TM"}
        if any(p not in allowed for p in polarisations):
            raise ValueError(f"polarisations must be drawn from {allowed}.")
        beta_path = np.asarray(beta_path)
        if beta_path.shape[0] != 2:
            raise ValueError("beta_path must be a 2×N array of Bloch vectors.")
        ns = beta_path.shape[1]
        n = self.Nx * self.Ny
        frequencies = {p: np.zeros((num_bands, ns), dtype=float) for p in polarisations}
        eigenvalues = {p: np.zeros((num_bands, ns), dtype=complex) for p in polarisations}
        blas = _eigs.DeviceBlas(self.device)
        dense = inner == "lu" or (inner == "auto" and n <= self.DENSE_LIMIT)
        # effective shift: left of the (non-negative) spectrum when the requested shift is 0
        # The reference's default shift sigma = 0 sits ON the spectrum at Gamma (A is singular there) and ARPACK
        # survives it only through SuperLU's perturbed pivots; the shift actually used for sigma = 0 is -delta
        # (left of the non-negative spectrum: same "k nearest" set).  A caller's non-zero shift is used as given,
        # so that its nearest-eigenvalue set is the reference's eigs(sigma=eig_sigma).
        delta = (2 * np.pi * 0.05 / self.a) ** 2
        if eig_sigma == 0:
            sigma = -delta
        else:
            sigma = float(np.real(eig_sigma)) if np.imag(eig_sigma) == 0 else complex(eig_sigma)
        info = dict(restarts=0, inner="lu" if dense else "krylov", eigensolves=0, max_residual=0.0)
        for pol in ("TM", "TE"):          # reference order inside the loop (:311-323)
            if pol not in polarisations:
                continue
            kind, ca, cb, M = self._pencil_arrays(pol)
            A0 = self._operator(kind, ca, cb, None, (1, 1))
            dev, tdt = A0.device, A0.tdtype
            Mdev = torch.from_numpy(M).to(dev).to(tdt).reshape(-1)
            cdev = (-sigma) * Mdev.reshape(self.Ny, self.Nx)
            if inner in ("auto", "block") and self._block_path_ok(ca, cb, M, sigma, num_bands, inner == "block"):
                info["inner"] = "block-lu"
                self._sweep_device(kind, A0, Mdev, beta_path, num_bands, sigma, tol, delta, pol, eigenvalues,
                                   frequencies, info)
                continue
            if dense:
                self._sweep_batched(kind, A0, cdev, Mdev, beta_path, num_bands, sigma, tol, delta, pol,
                                    eigenvalues, frequencies, info, batch)
                continue
            v0 = None
            for idx in range(ns):
                ph = self._phases(beta_path[:, idx])
                A = self._operator(kind, A0.ca, A0.cb, None, ph)
                As = self._operator(kind, A0.ca, A0.cb, cdev, ph)
                solve = lambda r: As.solve(r, method="bicgstab", precond="jacobi", rtol=1e-13,  # noqa: E731
                                           maxit=20000)[0]
                lam, X, res, rs = _eigs.eigs_shift_invert(A.apply, lambda v: Mdev * v.reshape(-1), solve, n,
                                                           num_bands, sigma, tol=tol, v0=v0, device=dev, dtype=tdt,
                                                           blas=blas, seed=idx)
                v0 = X.sum(dim=0)          # warm start for the next path point
                order = np.argsort(np.real(lam))
                vals = lam[order][:num_bands]
                eigenvalues[pol][:, idx] = vals
                frequencies[pol][:, idx] = self._normalise_eigenvalues(vals)
                info["restarts"] += rs
                info["eigensolves"] += 1
                info["max_residual"] = max(info["max_residual"], float(np.max(res[np.abs(lam) > 1e-6 * delta], initial=0.0)))
                A.close()
                As.close()
        self.solver_info = info
        result = BandStructureResult(beta_path=beta_path, tick_positions=self._tick_positions_from_path(beta_path),
                                     tick_labels=getattr(self, "_tick_labels", []) or
                                     [""] * len(self._tick_positions_from_path(beta_path)),
                                     frequencies=frequencies, eigenvalues=eigenvalues)
        self._results = result
        return result

    def _block_path_ok(self, ca, cb, M, sigma, num_bands, explicit):
        """The one-CTA-per-path-point kernels (csrc/band.cu) need Bloch walls on both axes, a cell whose block
        factorisation fits in shared memory and a Hermitian positive definite shifted operator (real positive
        materials, shift left of the spectrum)."""
        L = _lib.lib()
        ok = (tuple(int(b) for b in self.boundary_conditions) == (1, 1) and 3 <= self.Nx <= L.fdfd_band_max_nx()
              and self.Ny >= 3 and (3 * self.Nx * self.Ny + 2 * self.Nx + 41) * 16 <= 220 * 1024
              and np.imag(sigma) == 0 and np.real(sigma) < 0 and 2 * num_bands + 1 <= 40
              and all(np.all(np.imag(a) == 0) and np.all(np.real(a) > 0) for a in (ca, cb, M)))
        if explicit and not ok:
            raise ValueError("inner='block' needs Bloch walls on both axes, real positive materials, a shift left of "
                             "the spectrum and a cell that fits the shared-memory block factorisation")
        return ok

    @staticmethod
    def _projected_chunk(Hm, b, k, keep):
        """Ritz pairs, residual estimates and the thick-restart basis of a stack of projected matrices."""
        ncv = Hm.shape[-1]
        mu, S = np.linalg.eig(Hm)
        order = np.argsort(-np.abs(mu), axis=1)
        mu = np.take_along_axis(mu, order, axis=1)
        S = np.take_along_axis(S, order[:, None, :], axis=2)
        S = S / np.linalg.norm(S, axis=1, keepdims=True)
        res = np.abs(np.einsum("pc,pck->pk", b, S))
        Sk = np.ascontiguousarray(S[:, :, :k].transpose(0, 2, 1))
        Q, _ = np.linalg.qr(S[:, :, :keep])
        Tn = np.zeros((Hm.shape[0], ncv + 1, ncv), dtype=np.complex128)
        Tn[:, :keep, :keep] = np.conj(Q.transpose(0, 2, 1)) @ Hm @ Q
        Tn[:, keep, :keep] = np.einsum("pc,pck->pk", b, Q)
        return mu, res, Sk, np.ascontiguousarray(Q.transpose(0, 2, 1)), Tn

    @classmethod
    def _projected_problems(cls, Hm, b, k, keep, min_chunk=24):
        """`_projected_chunk` over all active path points; the points are independent, so the stack is cut into
        chunks that run on a thread pool (LAPACK releases the GIL) — the result does not depend on the cut."""
        import os
        from concurrent.futures import ThreadPoolExecutor
        P = Hm.shape[0]
        workers = max(1, min(8, (os.cpu_count() or 1), P // min_chunk))
        if workers == 1:
            return cls._projected_chunk(Hm, b, k, keep)
        cuts = np.array_split(np.arange(P), workers)
        with ThreadPoolExecutor(max_workers=workers) as pool:
            parts = list(pool.map(lambda c: cls._projected_chunk(Hm[c], b[c], k, keep), cuts))
        return tuple(np.concatenate([q[i] for q in parts], axis=0) for i in range(5))

    def _sweep_device(self, kind, A0, Mdev, beta_path, num_bands, sigma, tol, delta, pol, eigenvalues, frequencies,
                      info, max_restarts=100, seed=0):
        """All path points of one polarisation in lock-step on this library's own kernels (csrc/band.cu, C-ABI
        fdfd_band_*): block LU of the cyclic block-tridiagonal shifted operator, Arnoldi cycles and Krylov-Schur
        restarts on the device, one CTA per path point; the host only decomposes the (ncv+1) x ncv projected
        matrices (same restart logic as `eigs.krylov_schur_batched`)."""
        import ctypes as C
        import time
        import torch
        L = _lib.lib()
        n, P, k = self.Nx * self.Ny, beta_path.shape[1], int(num_bands)
        dev, tdt = A0.device, A0.tdtype
        ncv = int(min(max(2 * k + 1, 20), n, 40))
        keep = min(ncv - 1, k + max(1, (ncv - k) // 2))
        stream = torch.cuda.current_stream(dev).cuda_stream

        def lap(key, t0):
            torch.cuda.synchronize(dev)
            info[key] = info.get(key, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        t = time.perf_counter()
        ph = np.array([self._phases(beta_path[:, i]) for i in range(P)])
        phases = np.ascontiguousarray(np.stack([ph[:, 0].real, ph[:, 0].imag, ph[:, 1].real, ph[:, 1].imag], axis=1))
        Mc = Mdev.to(torch.complex128).contiguous()
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(L.fdfd_band_create(C.byref(h), {"TE": _lib.POL_TE, "TM": _lib.POL_TM}[kind], self.Nx, self.Ny, P,
                                          A0.ca.data_ptr(), A0.cb.data_ptr(), Mc.data_ptr(), A0.sx, A0.sy, float(sigma),
                                          phases.ctypes.data, ncv, k, stream))
        try:
            g = torch.Generator(device="cpu").manual_seed(int(seed))
            v0 = torch.view_as_complex(torch.rand((P, n, 2), generator=g, dtype=torch.float64) * 2 - 1).to(dev)
            v0 = (v0 / torch.linalg.vector_norm(v0, dim=1, keepdim=True)).contiguous()
            _lib.check(L.fdfd_band_set_start(h, v0.data_ptr(), stream))
            t = lap("t_factor", t)
            Hh = np.zeros((P, ncv + 1, ncv), dtype=np.complex128)
            active = np.arange(P, dtype=np.int32)
            lam_out = np.zeros((P, k), dtype=np.complex128)
            res_out = np.zeros((P, k))
            p0, restarts = 0, 0
            while True:
                _lib.check(L.fdfd_band_cycle(h, p0, ncv, Hh.ctypes.data, stream))
                t = lap("t_krylov", t)
                # projected problems of all active points at once (stacked LAPACK calls).  The operator is
                # self-adjoint in the M inner product, so its Ritz vectors are well conditioned and the thick
                # restart keeps the span of the `keep` dominant ones, orthonormalised by QR (the same subspace
                # an ordered Schur form would keep): A (V Q) = (V Q)(Q^H H Q) + v (b Q).
                mu, res, Sk, Qk, Tn = self._projected_problems(Hh[active, :ncv, :ncv], Hh[active, ncv, :ncv], k, keep)
                done = np.all(res[:, :k] <= tol * np.abs(mu[:, :k]), axis=1) | (restarts >= max_restarts)
                mus = mu[:, :k]
                t = lap("t_host_projected", t)
                fin = np.flatnonzero(done)
                if fin.size:
                    idx = np.ascontiguousarray(active[fin], dtype=np.int32)
                    lam = np.ascontiguousarray(sigma + 1.0 / mus[fin])
                    res = np.zeros((fin.size, k))
                    Sf = np.ascontiguousarray(Sk[fin])
                    _lib.check(L.fdfd_band_ritz(h, idx.ctypes.data, int(fin.size), Sf.ctypes.data, lam.ctypes.data, k,
                                                res.ctypes.data, stream))
                    lam_out[idx], res_out[idx] = lam, res
                go = np.flatnonzero(~done)
                t = lap("t_post", t)
                if go.size == 0:
                    break
                active = np.ascontiguousarray(active[go], dtype=np.int32)
                _lib.check(L.fdfd_band_set_active(h, active.ctypes.data, int(active.size), stream))
                Qg, Tg = np.ascontiguousarray(Qk[go]), np.ascontiguousarray(Tn[go])
                _lib.check(L.fdfd_band_restart(h, Qg.ctypes.data, Tg.ctypes.data, keep, stream))
                p0 = keep
                restarts += 1
        finally:
            L.fdfd_band_destroy(h)
        for i in range(P):
            order = np.argsort(np.real(lam_out[i]))
            vals = lam_out[i][order][:k]
            eigenvalues[pol][:, i] = vals
            frequencies[pol][:, i] = self._normalise_eigenvalues(vals)
            info["max_residual"] = max(info["max_residual"],
                                       float(np.max(res_out[i][np.abs(lam_out[i]) > 1e-6 * delta], initial=0.0)))
        info["restarts"] += restarts
        info["eigensolves"] += P

    def _sweep_batched(self, kind, A0, cdev, Mdev, beta_path, num_bands, sigma, tol, delta, pol, eigenvalues,
                       frequencies, info, batch):
        """All path points of one polarisation in lock-step: the shifted operator
        is an affine function of the two Bloch phases, so the dense pieces are assembled once, each
        chunk of points forms its shifted matrices by broadcasting, inverts them as one batch and
        runs `eigs.krylov_schur_batched` — one batched GEMV per Arnoldi step for the whole chunk."""
        import torch
        import time
        n, ns = self.Nx * self.Ny, beta_path.shape[1]
        dev, tdt = A0.device, A0.tdtype

        def lap(key, t0):
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            info[key] = info.get(key, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        t = time.perf_counter()
        C0, Cx, Cxp, Cy, Cyp = self._dense_pieces(kind, A0.ca, A0.cb, cdev)
        t = lap("t_assemble", t)
        if batch is None:                     # S, its inverse and one temporary per point
            free = torch.cuda.mem_get_info(dev)[0] if dev.type == "cuda" else 8 << 30
            batch = max(1, min(ns, int(0.6 * free // (3 * n * n * 16))))
        ph = np.array([self._phases(beta_path[:, i]) for i in range(ns)])
        for lo in range(0, ns, batch):
            hi = min(ns, lo + batch)
            px = torch.from_numpy(ph[lo:hi, 0]).to(dev).to(tdt).view(-1, 1, 1)
            py = torch.from_numpy(ph[lo:hi, 1]).to(dev).to(tdt).view(-1, 1, 1)
            S = C0.unsqueeze(0) + px * Cx
            S += px.conj() * Cxp
            S += py * Cy
            S += py.conj() * Cyp
            inv = torch.linalg.inv(S)
            t = lap("t_invert", t)
            state = {"inv": inv}

            def set_active(idx, state=state, inv=inv):
                state["inv"] = inv if idx is None else inv.index_select(0, idx)

            def op(W, state=state):
                return torch.bmm(state["inv"], (Mdev * W).unsqueeze(2)).squeeze(2)

            mu, X, _, rs = _eigs.krylov_schur_batched(op, set_active, hi - lo, n, num_bands, tol=tol, device=dev,
                                                       dtype=tdt, seed=lo)
            t = lap("t_krylov", t)
            lam = sigma + 1.0 / mu                                             # (P, k)
            lam_d = torch.from_numpy(lam).to(dev).to(tdt)
            MX = Mdev * X                                                     # (P, k, n)
            AX = torch.bmm(X, S.transpose(1, 2)) + sigma * MX                 # A = S + sigma M
            num = torch.linalg.vector_norm(AX - lam_d.unsqueeze(2) * MX, dim=2)
            den = torch.linalg.vector_norm(AX, dim=2) + lam_d.abs() * torch.linalg.vector_norm(MX, dim=2)
            res = (num / den.clamp_min(1e-300)).cpu().numpy()
            for i in range(hi - lo):
                order = np.argsort(np.real(lam[i]))
                vals = lam[i][order][:num_bands]
                eigenvalues[pol][:, lo + i] = vals
                frequencies[pol][:, lo + i] = self._normalise_eigenvalues(vals)
                info["max_residual"] = max(info["max_residual"],
                                           float(np.max(res[i][np.abs(lam[i]) > 1e-6 * delta], initial=0.0)))
            info["restarts"] += rs
            info["eigensolves"] += hi - lo
            del S, inv, state
            t = lap("t_post", t)

    def _sort_eigenvalues(self, values, num_bands):
        return values[np.argsort(np.real(values))][:num_bands]

    def _normalise_eigenvalues(self, values):
        vals = np.clip(np.real_if_close(values).real, 0.0, None)
        return self.a / (2 * np.pi) * np.sqrt(vals)

    # ---- plotting: optional host-side matplotlib (Band_Diagram_Solver.py:357-531) ---------------------
    def plot_band_diagram(self, result=None, *, wnmax=None, ax=None, show=True):
        import matplotlib.pyplot as plt
        result = result or self._results
        if result is None:
            raise RuntimeError("Run compute_band_structure() first")
        if ax is None:
            _, ax = plt.subplots(figsize=(7, 5))
        xs = np.arange(result.beta_path.shape[1])
        for pol, style in (("TM", "b-"), ("TE", "r--")):
            if pol in result.frequencies:
                for band in result.frequencies[pol]:
                    ax.plot(xs, band, style, lw=1)
        ax.set_xticks(result.tick_positions)
        ax.set_xticklabels(result.tick_labels)
        ax.set_ylabel("a/λ")
        if wnmax is not None:
            ax.set_ylim(0, wnmax)
        if show:
            plt.show()
        return ax
