"""GPU drop-in for the reference's ``PeriodicModeSolver3D``
(Periodic_Solver_3D/Periodic_Solver_3D.py:17-963): unit cell periodic in z, four transverse
fields, generalised pencil ``A x = gamma B x``.  Same constructor, ``add_block`` / ``add_pec`` /
``add_pmc`` / ``add_UPML``, ``solve(method="refined"|"eigs", ...)``, ``store_fields``,
``save_results`` / ``load_results`` (NPZ layout unchanged), attributes ``eigenvalues``,
``eigenvectors``, ``gammas``, ``refined_residuals``, ``refined_restarts``, ``fields``.

GPU side: the pencil matrices live on the device as CSR (K2b), the refined shift-invert Arnoldi
(``fdfd_b200.eigs.refined_shift_invert_arnoldi``) keeps its basis, Gram-Schmidt, SpMM and refined
extraction on the device; ``method="eigs"`` uses the Krylov-Schur solver (K6).  The shifted solve
is a direct solver on the device: the z-periodic cell is a ring of planes, folded into a block
tridiagonal matrix and factorised with dense block pivots (``eigs.PlaneBlockShiftedSolver``); cells
below 4000 unknowns use one dense LU (``eigs.DenseShiftedSolver``).  Nothing is sent to a CPU path.
Host side (numpy, as the reference): sub-pixel blocks, penalty PEC/PMC, per-layer UPML,
cell -> Yee averaging, operator and pencil assembly (vectorised; arrays identical).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import eigs as _eigs

_COMPS = ("xx", "yy", "zz")


class PeriodicModeSolver3D:
    DENSE_LIMIT = 45000          # dense-LU shifted solve (small cells, or PLANE_BLOCK = False)
    PLANE_BLOCK = True           # plane-block direct solver (eigs.PlaneBlockShiftedSolver) for cells >= 4000 unknowns

    def __init__(self, Nx, Ny, Nz, x_range, y_range, z_range, freq, num_modes, sigma_guess=None, tol=0, ncv=None,
                 *, device=None):
        self.Nx, self.Ny, self.Nz = int(Nx), int(Ny), int(Nz)
        self.x_range, self.y_range, self.z_range = x_range, y_range, z_range
        self.dx, self.dy, self.dz = x_range / self.Nx, y_range / self.Ny, z_range / self.Nz
        self.N = self.Nx * self.Ny * self.Nz
        self.freq = freq
        self.omega = 2 * np.pi * freq
        self.k0 = self.omega / 3e8                # Periodic_Solver_3D.py:33-35 (the reference's own constants)
        self.epsilon0 = 8.85e-12
        self.mu0 = 1.26e-6
        self.num_modes = int(num_modes)
        self.sigma_guess = sigma_guess if sigma_guess is not None else 0
        self.tol, self.ncv = tol, ncv
        nx, ny, nz = self.Nx, self.Ny, self.Nz
        self.shape_cell = self.shape_hz = (nx, ny, nz)
        self.shape_ex = self.shape_hy = (nx, ny + 1, nz)
        self.shape_ey = self.shape_hx = (nx + 1, ny, nz)
        self.shape_ez = (nx + 1, ny + 1, nz)
        self.n_ex = self.n_hy = int(np.prod(self.shape_ex))
        self.n_ey = self.n_hx = int(np.prod(self.shape_ey))
        self.n_ez = int(np.prod(self.shape_ez))
        self.n_hz = self.N
        for name in ("Erxx", "Eryy", "Erzz", "Mrxx", "Mryy", "Mrzz"):
            setattr(self, f"cell_{name}_3D", np.ones(self.shape_cell, dtype=complex))
        self.material_no_average_mask = np.zeros(self.shape_cell, dtype=bool)
        self._pec_regions, self._pmc_regions = [], []
        self.device = device
        self.update_component_materials()
        self.fields = {}
        self.eigenvalues = self.eigenvectors = None
        self.solver_info = {}

    def _invalidate_solution(self):
        self.fields = {}
        self.eigenvalues = self.eigenvectors = None
        self.__dict__.pop("gammas", None)

    # ---- operators (Periodic_Solver_3D.py:95-194): GPU assembly ---------------------------------------------
    _OPERATORS = ("DEX_EZ_TO_EX", "DEY_EZ_TO_EY", "DEX_EY_TO_HZ", "DEY_EX_TO_HZ", "DHX_HY_TO_EZ", "DHY_HX_TO_EZ",
                  "DHX_HZ_TO_HX", "DHY_HZ_TO_HY", "DEZ_EX", "DEZ_EY", "DHZ_HX", "DHZ_HY", "AZ_EX", "AZ_EY", "AZ_HX", "AZ_HY")

    def __getattr__(self, name):
        # the sixteen staggered operators are assembled on the GPU on first use (the reference builds them
        # with Python triple loops in its constructor, Periodic_Solver_3D.py:95-194)
        if name in type(self)._OPERATORS and "shape_ex" in self.__dict__:
            self._init_operators()
            return self.__dict__[name]
        raise AttributeError(name)

    def _init_operators(self):
        """K1b kernel (csrc/yee_assemble.cu: stagger_fill_kernel) through fdfd_b200.stagger: two entries per
        row, byte-identical to the reference's loops + coo -> csr; H operators are the CSC reading of the same
        index arrays with negated data (-D^H)."""
        from . import stagger
        d = stagger.difference
        self.DEX_EZ_TO_EX, self.DHX_HY_TO_EZ = d(self.shape_ez, self.shape_ex, self.dx, 0, with_h=True)
        self.DEY_EZ_TO_EY, self.DHY_HX_TO_EZ = d(self.shape_ez, self.shape_ey, self.dy, 1, with_h=True)
        self.DEX_EY_TO_HZ, self.DHX_HZ_TO_HX = d(self.shape_ey, self.shape_hz, self.dx, 0, with_h=True)
        self.DEY_EX_TO_HZ, self.DHY_HZ_TO_HY = d(self.shape_ex, self.shape_hz, self.dy, 1, with_h=True)
        # periodic z difference: forward weight 1/dz, self weight -1/dz (no Bloch phase, :172)
        self.DEZ_EX, self.DHZ_HY = stagger.stagger_csr(self.shape_ex, self.shape_ex, 2, -1.0 / self.dz, 1.0 / self.dz, with_h=True)
        self.DEZ_EY, self.DHZ_HX = stagger.stagger_csr(self.shape_ey, self.shape_ey, 2, -1.0 / self.dz, 1.0 / self.dz, with_h=True)
        self.AZ_EX = stagger.periodic_z(self.shape_ex, 0.5, 0.5)
        self.AZ_EY = stagger.periodic_z(self.shape_ey, 0.5, 0.5)
        self.AZ_HX = self.AZ_EY.conj().T
        self.AZ_HY = self.AZ_EX.conj().T

    # ---- modelling helpers (Periodic_Solver_3D.py:196-392) --------------------------------------------------
    @staticmethod
    def _three(name, value):
        if np.isscalar(value):
            return np.full(3, value, dtype=complex)
        arr = np.asarray(value, dtype=complex)
        if arr.ndim == 1 and arr.size == 3:
            return arr
        raise ValueError(f"{name} must be a scalar or a length-3 1D array (xx, yy, zz).")

    def _step(self, axis):
        return {"x": self.dx, "y": self.dy, "z": self.dz}[axis]

    def _to_length(self, value, axis):
        if isinstance(value, (int, np.integer)):
            return int(value) * self._step(axis)
        if isinstance(value, (float, np.floating)):
            return float(value)
        raise ValueError("Coordinates must be int grid indices or float physical positions in metres.")

    def _span(self, name, values, axis):
        if isinstance(values, slice):
            if values.step not in (None, 1):
                raise ValueError(f"{name} slice step must be None or 1.")
            values = (0 if values.start is None else values.start,
                      {"x": self.Nx, "y": self.Ny, "z": self.Nz}[axis] if values.stop is None else values.stop)
        try:
            lo, hi = values
        except (TypeError, ValueError):
            raise ValueError(f"{name} must be a slice or a (min, max) pair.")
        lo, hi = self._to_length(lo, axis), self._to_length(hi, axis)
        if hi <= lo:
            raise ValueError(f"{name} must satisfy max > min.")
        if lo < 0 or hi > {"x": self.x_range, "y": self.y_range, "z": self.z_range}[axis]:
            raise ValueError(f"{name} is out of bounds of the simulation grid.")
        return lo, hi

    def _bbox(self, spans):
        out = []
        for (lo, hi), axis, N in zip(spans, "xyz", (self.Nx, self.Ny, self.Nz)):
            out += [max(0, int(np.floor(lo / self._step(axis)))), min(N, int(np.ceil(hi / self._step(axis))))]
        return out

    def _region_slices(self, x_range, y_range, z_range):
        spans = [self._span(n, r, a) for n, r, a in (("x_range", x_range, "x"), ("y_range", y_range, "y"),
                                                     ("z_range", z_range, "z"))]
        x0, x1, y0, y1, z0, z1 = self._bbox(spans)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            raise ValueError("Region is outside the simulation grid.")
        return slice(x0, x1), slice(y0, y1), slice(z0, z1)

    def _cell(self, prefix, comp):
        if prefix not in ("eps", "mu"):
            raise ValueError(f"Unknown {prefix} component {comp!r}.")
        return getattr(self, f"cell_{'Er' if prefix == 'eps' else 'Mr'}{comp}_3D")

    @staticmethod
    def _subpixel_axis(start, stop, step, sub):
        idx = np.arange(start, stop, dtype=float)
        off = (np.arange(sub, dtype=float) + 0.5) / sub
        return (idx[:, None] + off[None, :]) * step

    def add_block(self, er, mr, x_range, y_range, z_range, *, subpixels=8):
        spans = [self._span(n, r, a) for n, r, a in (("x_range", x_range, "x"), ("y_range", y_range, "y"),
                                                     ("z_range", z_range, "z"))]
        subpixels = int(subpixels)
        if subpixels <= 0:
            raise ValueError("subpixels must be positive.")
        x0, x1, y0, y1, z0, z1 = self._bbox(spans)
        if x0 >= x1 or y0 >= y1 or z0 >= z1:
            return
        ins = []
        for (lo, hi), (a0, a1), axis in zip(spans, ((x0, x1), (y0, y1), (z0, z1)), "xyz"):
            pts = self._subpixel_axis(a0, a1, self._step(axis), subpixels)
            ins.append((pts >= lo) & (pts <= hi))
        fraction = (ins[0][:, :, None, None, None, None] & ins[1][None, None, :, :, None, None]
                    & ins[2][None, None, None, None, :, :]).mean(axis=(1, 3, 5))
        sl = (slice(x0, x1), slice(y0, y1), slice(z0, z1))
        er, mr = self._three("er", er), self._three("mr", mr)
        hit = (fraction > 0.0) & ~self.material_no_average_mask[sl]
        if not np.any(hit):
            return
        for prefix, vals in (("eps", er), ("mu", mr)):
            for comp, v in zip(_COMPS, vals):
                tgt = self._cell(prefix, comp)[sl]
                tgt[hit] = tgt[hit] * (1.0 - fraction[hit]) + v * fraction[hit]
        self.update_component_materials()
        self._invalidate_solution()

    def _penalty(self, prefix, regions, x_range, y_range, z_range, components, value):
        sl = self._region_slices(x_range, y_range, z_range)
        for comp in (_COMPS if components is None else tuple(components)):
            self._cell(prefix, comp)[sl] = value
        self.material_no_average_mask[sl] = True
        regions.append(sl)
        self.update_component_materials()
        self._invalidate_solution()

    def add_pec(self, x_range, y_range, z_range, components=None, epsilon=1e8):
        self._penalty("eps", self._pec_regions, x_range, y_range, z_range, components, epsilon)

    def add_pmc(self, x_range, y_range, z_range, components=None, mu=1e8):
        self._penalty("mu", self._pmc_regions, x_range, y_range, z_range, components, mu)

    def add_UPML(self, sides=('-x', '+x', '-y', '+y'), width=10, max_loss=5, n=3):
        """Per-layer in-place scaling exactly as the reference (Periodic_Solver_3D.py:341-392), including
        its naming quirk: '+y' stretches the LOW-y layers and '-y' the HIGH-y layers."""
        assert width > 0
        assert width <= self.Nx // 2 and width <= self.Ny // 2, "PML width too large for grid."
        where = {'-x': lambda i: np.s_[i, :, :], '+x': lambda i: np.s_[-1 - i, :, :],
                 '+y': lambda i: np.s_[:, i, :], '-y': lambda i: np.s_[:, -1 - i, :]}
        for i in range(width):
            sigma = max_loss * ((width - i) / width) ** n
            S = 1 + 1j * sigma / (self.omega * self.epsilon0)
            for side in ('-x', '+x', '+y', '-y'):
                if side not in sides:
                    continue
                sl = where[side](i)
                normal = "xx" if side[1] == "x" else "yy"
                for pre in ("Er", "Mr"):
                    for comp in _COMPS:
                        arr = getattr(self, f"cell_{pre}{comp}_3D")
                        if comp == normal:
                            arr[sl] /= S
                        else:
                            arr[sl] *= S
        self.update_component_materials()
        self._invalidate_solution()

    # ---- cell -> Yee averaging (Periodic_Solver_3D.py:398-475) -----------------------------------------------
    def _average(self, values, offsets, keep):
        nx, ny, nz = self.shape_cell
        shape = (nx + max(o[0] for o in offsets), ny + max(o[1] for o in offsets), nz)
        acc = np.zeros(shape, dtype=complex)
        cnt = np.zeros(shape, dtype=float)
        for ox, oy in offsets:
            acc[ox:ox + nx, oy:oy + ny, :] += values
            cnt[ox:ox + nx, oy:oy + ny, :] += 1
        acc = acc / cnt
        if keep is not None:
            ii, jj, kk = np.nonzero(keep)
            for ox, oy in offsets:
                acc[ii + ox, jj + oy, kk] = values[ii, jj, kk]
        return acc

    _X = ((0, 0), (1, 0))
    _Y = ((0, 0), (0, 1))
    _XY = ((0, 0), (1, 0), (0, 1), (1, 1))

    def _material_on_fields(self, erxx, eryy, erzz, mrxx, mryy, mrzz, keep):
        return {"erxx": self._average(erxx, self._Y, keep), "eryy": self._average(eryy, self._X, keep),
                "erzz": self._average(erzz, self._XY, keep), "mrxx": self._average(mrxx, self._X, keep),
                "mryy": self._average(mryy, self._Y, keep), "mrzz": mrzz.copy()}

    def update_component_materials(self):
        mats = self._material_on_fields(self.cell_Erxx_3D, self.cell_Eryy_3D, self.cell_Erzz_3D, self.cell_Mrxx_3D,
                                        self.cell_Mryy_3D, self.cell_Mrzz_3D, self.material_no_average_mask)
        for k, v in mats.items():
            setattr(self, f"{k[0].upper()}{k[1:]}_3D", v.copy())
        return mats

    # ---- pencil (Periodic_Solver_3D.py:478-533) ---------------------------------------------------------------
    def _build_eigen_matrices(self):
        w, e0, m0 = self.omega, self.epsilon0, self.mu0
        dg = lambda a: sp.diags(a.ravel(order="F"), format="csr")  # noqa: E731
        Erxx, Eryy, Erzz, Mrxx, Mryy, Mrzz = (dg(getattr(self, f"{n}_3D")) for n in
                                              ("Erxx", "Eryy", "Erzz", "Mrxx", "Mryy", "Mrzz"))
        Z = lambda r, c: sp.csr_matrix((r, c), dtype=complex)  # noqa: E731
        ezi, mzi = Erzz.power(-1), Mrzz.power(-1)
        A = sp.bmat([
            [self.DEZ_EX, Z(self.n_ex, self.n_ey),
             self.DEX_EZ_TO_EX @ (-1j / (w * e0) * ezi @ self.DHY_HX_TO_EZ),
             self.DEX_EZ_TO_EX @ (1j / (w * e0) * ezi @ self.DHX_HY_TO_EZ) + 1j * w * m0 * Mryy],
            [Z(self.n_ey, self.n_ex), self.DEZ_EY,
             self.DEY_EZ_TO_EY @ (-1j / (w * e0) * ezi @ self.DHY_HX_TO_EZ) - 1j * w * m0 * Mrxx,
             self.DEY_EZ_TO_EY @ (1j / (w * e0) * ezi @ self.DHX_HY_TO_EZ)],
            [self.DHX_HZ_TO_HX @ (1j / (w * m0) * mzi @ self.DEY_EX_TO_HZ),
             self.DHX_HZ_TO_HX @ (-1j / (w * m0) * mzi @ self.DEX_EY_TO_HZ) - 1j * w * e0 * Eryy,
             self.DHZ_HX, Z(self.n_hx, self.n_hy)],
            [self.DHY_HZ_TO_HY @ (1j / (w * m0) * mzi @ self.DEY_EX_TO_HZ) + 1j * w * e0 * Erxx,
             self.DHY_HZ_TO_HY @ (-1j / (w * m0) * mzi @ self.DEX_EY_TO_HZ), Z(self.n_hy, self.n_hx), self.DHZ_HY],
        ]).tocsr()
        B = sp.bmat([[self.AZ_EX, Z(self.n_ex, self.n_ey), None, None], [Z(self.n_ey, self.n_ex), self.AZ_EY, None, None],
                     [None, None, self.AZ_HX, Z(self.n_hx, self.n_hy)], [None, None, Z(self.n_hy, self.n_hx), self.AZ_HY]],
                    format="csr")
        return A, B

    def _unknown_planes(self):
        """z-plane index of every unknown of [Ex; Ey; Hx; Hy] (flat index i + nx (j + ny k), k slowest)."""
        return np.concatenate([np.arange(int(np.prod(sh))) // (sh[0] * sh[1])
                               for sh in (self.shape_ex, self.shape_ey, self.shape_hx, self.shape_hy)])

    # ---- solve (Periodic_Solver_3D.py:535-604) -----------------------------------------------------------------
    def solve(self, method="refined", sigma_guess=None, tol=None, ncv=None, max_restarts=12, random_seed=0):
        import torch
        from .operators import CsrOperator

        sigma_guess = self.sigma_guess if sigma_guess is None else sigma_guess
        tol = self.tol if tol is None else tol
        ncv = self.ncv if ncv is None else ncv
        if method not in ("eigs", "refined"):
            raise ValueError("method must be 'eigs' or 'refined'.")
        A, B = self._build_eigen_matrices()
        dev = torch.device(self.device or "cuda")
        planes = self._unknown_planes() if self.PLANE_BLOCK else None
        if method == "refined":
            (self.eigenvalues, self.eigenvectors, self.refined_residuals,
             self.refined_restarts) = _eigs.refined_shift_invert_arnoldi(
                A, B, sigma=sigma_guess, num_modes=self.num_modes, tol=tol, ncv=ncv, max_restarts=max_restarts,
                random_seed=random_seed, device=dev, dense_limit=self.DENSE_LIMIT, planes=planes, nplanes=self.Nz)
        else:
            Ad, Bd = CsrOperator(A, device=dev), CsrOperator(B, device=dev)
            solve = _eigs.make_shifted_solver(A, B, sigma_guess, dev, dense_limit=self.DENSE_LIMIT, planes=planes,
                                              nplanes=self.Nz)
            lam, X, res, _ = _eigs.eigs_shift_invert(Ad.apply, Bd.apply, solve, A.shape[0], self.num_modes, sigma_guess,
                                                      ncv=ncv, tol=max(float(tol or 0.0), 1e-12), device=dev)
            order = np.argsort(np.abs(lam - sigma_guess))
            self.eigenvalues = lam[order]
            self.eigenvectors = X.T.cpu().numpy()[:, order]
            self.refined_residuals = None
            self.refined_restarts = None
        self.gammas = self.eigenvalues / self.k0
        self.store_fields()

    def store_fields(self):
        edges = np.cumsum([0, self.n_ex, self.n_ey, self.n_hx, self.n_hy])
        for name, lo, hi, shape in zip(("Ex", "Ey", "Hx", "Hy"), edges[:-1], edges[1:],
                                       (self.shape_ex, self.shape_ey, self.shape_hx, self.shape_hy)):
            self.fields[name] = np.array([self.eigenvectors[lo:hi, m].reshape(shape, order="F")
                                          for m in range(self.num_modes)])

    # ---- result export (Periodic_Solver_3D.py:841-963): NPZ layout unchanged -----------------------------------
    def _results_dict(self, include_eigenvectors=False):
        if self.eigenvalues is None or not self.fields:
            raise RuntimeError("No results to store yet. Run solve() first.")
        out = dict(Nx=np.int64(self.Nx), Ny=np.int64(self.Ny), Nz=np.int64(self.Nz),
                   x_range=np.float64(self.dx * self.Nx), y_range=np.float64(self.dy * self.Ny),
                   z_range=np.float64(self.dz * self.Nz), freq=np.float64(self.freq),
                   num_modes=np.int64(self.num_modes), epsilon0=np.float64(self.epsilon0), mu0=np.float64(self.mu0),
                   sigma_guess=np.complex128(self.sigma_guess), tol=np.float64(self.tol),
                   material_no_average_mask=self.material_no_average_mask, eigenvalues=self.eigenvalues,
                   gammas=self.gammas, Ex=self.fields['Ex'], Ey=self.fields['Ey'], Hx=self.fields['Hx'],
                   Hy=self.fields['Hy'])
        for name in ("Erxx", "Eryy", "Erzz", "Mrxx", "Mryy", "Mrzz"):
            out[f"cell_{name}_3D"] = getattr(self, f"cell_{name}_3D")
        if include_eigenvectors and self.eigenvectors is not None:
            out['eigenvectors'] = self.eigenvectors
        if getattr(self, 'refined_residuals', None) is not None:
            out['refined_residuals'] = self.refined_residuals
            out['refined_restarts'] = np.int64(self.refined_restarts)
        return out

    def save_results(self, path, include_eigenvectors=False, compressed=True):
        (np.savez_compressed if compressed else np.savez)(path, **self._results_dict(include_eigenvectors))
        return path

    @classmethod
    def load_results(cls, path):
        d = np.load(path, allow_pickle=False)
        s = cls(int(d["Nx"]), int(d["Ny"]), int(d["Nz"]), float(d["x_range"]), float(d["y_range"]),
                float(d["z_range"]), float(d["freq"]), int(d["num_modes"]), sigma_guess=complex(d["sigma_guess"]),
                tol=float(d["tol"]))
        for name in ("Erxx", "Eryy", "Erzz", "Mrxx", "Mryy", "Mrzz"):
            setattr(s, f"cell_{name}_3D", d[f"cell_{name}_3D"])
        s.material_no_average_mask = d["material_no_average_mask"]
        s.update_component_materials()
        s.eigenvalues, s.gammas = d["eigenvalues"], d["gammas"]
        s.fields = {k: d[k] for k in ("Ex", "Ey", "Hx", "Hy")}
        if "eigenvectors" in d:
            s.eigenvectors = d["eigenvectors"]
        if "refined_residuals" in d:
            s.refined_residuals, s.refined_restarts = d["refined_residuals"], int(d["refined_restarts"])
        return s
