"""GPU drop-in for the reference's ``PeriodicModeSolver2D``
(Periodic_Solver_2D/Periodic_Solver_2D.py:12-590): TE / TM modes of a cell that is periodic in z, two fields
per polarisation ([Ex; Hy] for TM, [Hx; Ey] for TE), generalised pencil ``A x = gamma B x``.  Same constructor,
``add_rectangle`` / ``add_pec`` / ``add_pmc`` / ``add_pml``, ``solve(guess, tol, ncv)`` and result attributes
(``eigenvalues``, ``eigenvectors``, ``neff``, ``propagation_constant``, ``attenuation_constant``, ``Ex`` / ``Hy`` or
``Hx`` / ``Ey``, the candidate-index bookkeeping).

GPU side: the staggered operators come from the K1b assembly kernel (a (nx, nz) grid is the (nx, 1, nz) case of
the 3-D kernel), the pencil lives on the device as CSR (K2b), the eigenpairs nearest ``guess`` are computed by the
Krylov-Schur shift-invert solver (K6) whose shifted solve is the plane-block direct solver over the ring of z planes
(``eigs.PlaneBlockShiftedSolver``; cells below 4000 unknowns: one dense LU).  Host side (numpy, as the reference):
sub-pixel rectangles, penalty PEC / PMC, x-directed PML, cell -> field averaging, pencil assembly.

One deliberate difference (scripts/periodic2d_reproducibility.py): the reference hands the pencil to
ARPACK with ``tol=0`` and returns whatever it holds after the iteration limit — on its shipped example only the
first +- pair are eigenpairs of the pencil (relative residual 1e-11), the other six have residuals 1e-5 ... 3e-2
and move with the start vector.  This class returns the ``num_modes`` eigenpairs nearest ``guess`` converged to
``residuals`` <= 1e-9 (stored in ``solver_info``): it equals the reference wherever the reference converged.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import eigs as _eigs

_COMPS = ("xx", "yy", "zz")


class PeriodicModeSolver2D:
    DENSE_LIMIT = 45000
    PLANE_BLOCK = True

    def __init__(self, polarization, freq, x_range, z_range, Nx, Nz, num_modes, mode_filter=True, guess=0, tol=0,
                 ncv=None, *, device=None):
        polarization = str(polarization).upper()
        if polarization not in ("TE", "TM"):
            raise ValueError("polarization must be 'TE' or 'TM'.")
        self.polarization = polarization
        self.freq = self.frequency = freq
        self.x_range, self.z_range = x_range, z_range
        self.Nx, self.Nz = int(Nx), int(Nz)
        self.dx, self.dz = x_range / self.Nx, z_range / self.Nz
        self.N = self.Nx * self.Nz
        self.num_modes = int(num_modes)
        self.mode_filter = bool(mode_filter)
        self.epsilon0, self.mu0 = 8.85e-12, 1.26e-6          # the reference's own constants (:48-52)
        self.c = 1 / np.sqrt(self.epsilon0 * self.mu0)
        self.omega = 2 * np.pi * freq
        self.k0 = self.omega / self.c
        self.guess, self.tol, self.ncv = guess, tol, ncv
        nx, nz = self.Nx, self.Nz
        self.shape_cell = self.shape_ex = self.shape_hy = self.shape_hz = (nx, nz)
        self.shape_ez = self.shape_hx = self.shape_ey = (nx + 1, nz)
        self.n_ex = self.n_hy = nx * nz
        self.n_ez = self.n_hx = self.n_ey = (nx + 1) * nz
        self.n_hz = self.N
        for pre in ("eps", "mu"):
            for comp in _COMPS:
                setattr(self, f"cell_{pre}_r_{comp}", np.ones(self.shape_cell, dtype=complex))
        self.material_no_average_mask = np.zeros(self.shape_cell, dtype=bool)
        for pre in ("pec", "pmc"):
            for comp, shape in zip(_COMPS, ((self.shape_ex, self.shape_ey, self.shape_ez) if pre == "pec" else
                                           (self.shape_hx, self.shape_hy, self.shape_hz))):
                setattr(self, f"{pre}_{comp}_mask", np.zeros(shape, dtype=bool))
        self._pec_regions, self._pmc_regions = [], []
        self.device = device
        self.solver_info = {}
        self.update_component_materials()
        self._invalidate_solution()

    def _invalidate_solution(self):
        for name in ("eigenvalues", "eigenvectors", "neff", "propagation_constant", "attenuation_constant",
                     "spurious_scores", "accepted_candidate_indices", "rejected_candidate_indices",
                     "unselected_candidate_indices"):
            setattr(self, name, None)
        for name in ("Ex", "Ey", "Hx", "Hy"):
            self.__dict__.pop(name, None)

    # ---- operators (Periodic_Solver_2D.py:198-267): GPU assembly, on first use ------------------------------
    _OPERATORS = ("DEX_EZ_TO_EX", "DHX_HY_TO_EZ", "DEX_EY_TO_HZ", "DHX_HZ_TO_HX", "DEZ_EX", "DHZ_HY", "DEZ_EY", "DHZ_HX",
                  "AZ_EX", "AZ_HY", "AZ_EY", "AZ_HX")

    def __getattr__(self, name):
        if name in type(self)._OPERATORS and "shape_ex" in self.__dict__:
            self._init_operators()
            return self.__dict__[name]
        raise AttributeError(name)

    def _init_operators(self):
        """K1b kernel through fdfd_b200.stagger: the (nx, nz) grids are passed as (nx, 1, nz), x differences along
        axis 0, the periodic z difference / forward average along axis 2; two entries per row, byte-identical to the
        reference's loops + coo -> csr, H operators = the CSC reading of the same arrays with negated data."""
        from . import stagger
        three = lambda s: (s[0], 1, s[1])  # noqa: E731
        self.DEX_EZ_TO_EX, self.DHX_HY_TO_EZ = stagger.difference(three(self.shape_ez), three(self.shape_ex), self.dx, 0, with_h=True)
        self.DEX_EY_TO_HZ, self.DHX_HZ_TO_HX = stagger.difference(three(self.shape_ey), three(self.shape_hz), self.dx, 0, with_h=True)
        self.DEZ_EX, self.DHZ_HY = stagger.stagger_csr(three(self.shape_ex), three(self.shape_ex), 2, -1.0 / self.dz,
                                                       1.0 / self.dz, with_h=True)
        self.DEZ_EY, self.DHZ_HX = stagger.stagger_csr(three(self.shape_ey), three(self.shape_ey), 2, -1.0 / self.dz,
                                                       1.0 / self.dz, with_h=True)
        self.AZ_EX = stagger.periodic_z(three(self.shape_ex), 0.5, 0.5)
        self.AZ_EY = stagger.periodic_z(three(self.shape_ey), 0.5, 0.5)
        self.AZ_HY = self.AZ_EX.conj().T
        self.AZ_HX = self.AZ_EY.conj().T

    # ---- modelling helpers (Periodic_Solver_2D.py:115-196, 269-373) -----------------------------------------
    @staticmethod
    def _normalise_three(name, value):
        if np.isscalar(value):
            return np.full(3, value, dtype=complex)
        arr = np.asarray(value, dtype=complex)
        if arr.ndim == 1 and arr.size == 3:
            return arr
        raise ValueError(f"{name} must be a scalar or a length-3 1D array (xx, yy, zz).")

    @staticmethod
    def _validate_components(components):
        components = (components,) if isinstance(components, str) else tuple(components)
        bad = set(components) - set(_COMPS)
        if bad:
            raise ValueError(f"components contains invalid tensor component(s): {sorted(bad)}.")
        return components

    def _step(self, axis):
        return self.dx if axis == "x" else self.dz

    def _bound_to_index(self, value, axis):
        if isinstance(value, (int, np.integer)):
            return int(value)
        if isinstance(value, (float, np.floating)):
            return int(round(float(value) / self._step(axis)))
        raise ValueError("Region bounds must be int grid indices or float physical positions in metres.")

    def _coordinate_to_length(self, value, axis):
        if isinstance(value, (int, np.integer)):
            return int(value) * self._step(axis)
        if isinstance(value, (float, np.floating)):
            return float(value)
        raise ValueError("Coordinates must be int grid indices or float physical positions in metres.")

    def _range_to_lengths(self, name, values, axis):
        try:
            lo, hi = values
        except (TypeError, ValueError):
            raise ValueError(f"{name} must be a (min, max) pair.")
        lo, hi = self._coordinate_to_length(lo, axis), self._coordinate_to_length(hi, axis)
        if hi <= lo:
            raise ValueError(f"{name} must satisfy max > min.")
        if lo < 0 or hi > (self.x_range if axis == "x" else self.z_range):
            raise ValueError(f"{name} is out of bounds of the simulation grid.")
        return lo, hi

    def _region_slices(self, x_range, z_range):
        try:
            x0, x1 = (self._bound_to_index(v, "x") for v in (x_range[0], x_range[1]))
            z0, z1 = (self._bound_to_index(v, "z") for v in (z_range[0], z_range[1]))
        except (TypeError, IndexError):
            raise ValueError("x_range and z_range must be (min, max) pairs.")
        if not (x1 > x0 and z1 > z0):
            raise ValueError("x_range and z_range must satisfy max > min.")
        if not (0 <= x0 < x1 <= self.Nx and 0 <= z0 < z1 <= self.Nz):
            raise ValueError("Region is out of bounds of the simulation grid.")
        return slice(x0, x1), slice(z0, z1)

    def _cell_material_array(self, prefix, component):
        if prefix not in ("eps", "mu"):
            raise ValueError(f"Unknown {prefix} component {component!r}.")
        return getattr(self, f"cell_{prefix}_r_{component}")

    def _component_array(self, prefix, component):
        if prefix not in ("pec", "pmc"):
            raise ValueError(f"Unknown {prefix} component {component!r}.")
        return getattr(self, f"{prefix}_{component}_mask")

    @staticmethod
    def _subpixel_axis(start, stop, step, sub):
        idx = np.arange(start, stop, dtype=float)
        off = (np.arange(sub, dtype=float) + 0.5) / sub
        return (idx[:, None] + off[None, :]) * step

    def add_rectangle(self, epsilon, mu, x_range, z_range, *, subpixels=8):
        """Sub-pixel-smoothed rectangle on the cell grid (:303-317): volume fraction from `subpixels`^2 samples per
        cell, linear mixing with what the cell holds; cells frozen by add_pec / add_pmc are left alone."""
        spans = [self._range_to_lengths("x_range", x_range, "x"), self._range_to_lengths("z_range", z_range, "z")]
        subpixels = int(subpixels)
        if subpixels <= 0:
            raise ValueError("subpixels must be positive.")
        box = []
        for (lo, hi), axis, N in zip(spans, "xz", (self.Nx, self.Nz)):
            box.append((max(0, int(np.floor(lo / self._step(axis)))), min(N, int(np.ceil(hi / self._step(axis))))))
        if box[0][0] >= box[0][1] or box[1][0] >= box[1][1]:
            return
        inside = []
        for (lo, hi), (a0, a1), axis in zip(spans, box, "xz"):
            pts = self._subpixel_axis(a0, a1, self._step(axis), subpixels)
            inside.append((pts >= lo) & (pts <= hi))
        fraction = (inside[0][:, :, None, None] & inside[1][None, None, :, :]).mean(axis=(1, 3))
        sl = (slice(*box[0]), slice(*box[1]))
        epsilon, mu = self._normalise_three("epsilon", epsilon), self._normalise_three("mu", mu)
        hit = (fraction > 0.0) & ~self.material_no_average_mask[sl]
        if not np.any(hit):
            return
        for prefix, vals in (("eps", epsilon), ("mu", mu)):
            for comp, v in zip(_COMPS, vals):
                tgt = self._cell_material_array(prefix, comp)[sl]
                tgt[hit] = tgt[hit] * (1.0 - fraction[hit]) + v * fraction[hit]
        self.update_component_materials()
        self._invalidate_solution()

    def _penalty(self, prefix, regions, x_range, z_range, components, value):
        sl = self._region_slices(x_range, z_range)
        for comp in self._validate_components(_COMPS if components is None else components):
            self._cell_material_array(prefix, comp)[sl] = value
        self.material_no_average_mask[sl] = True
        regions.append(sl)
        self.update_component_materials()
        self._invalidate_solution()

    def add_pec(self, x_range, z_range, components=None, epsilon=1e8):
        """PEC-like region as a large-permittivity penalty (:319-334; no DOF elimination)."""
        self._penalty("eps", self._pec_regions, x_range, z_range, components, epsilon)

    def add_pmc(self, x_range, z_range, components=None, mu=1e8):
        self._penalty("mu", self._pmc_regions, x_range, z_range, components, mu)

    def add_pml(self, pml_width=30, n=3, sigma_max=5.0, direction="all"):
        """x-directed uniaxial PML (:348-373): Sx = 1 + j sigma_x / (eps0 omega); xx components divided, yy / zz
        multiplied, for eps and mu alike."""
        pml_width = int(pml_width)
        if pml_width <= 0:
            raise ValueError("pml_width must be positive.")
        if direction not in ("x-", "x+", "x", "all"):
            raise ValueError("direction must be one of 'x-', 'x+', 'x', or 'all'.")
        sigma_x = np.zeros(self.shape_cell, dtype=float)
        for i in range(min(pml_width, self.Nx)):
            s = sigma_max * ((pml_width - i) / pml_width) ** n
            if direction in ("x-", "x", "all"):
                sigma_x[i, :] = s
        for i in range(min(pml_width, self.Nx)):
            s = sigma_max * ((pml_width - i) / pml_width) ** n
            if direction in ("x+", "x", "all"):
                sigma_x[-i - 1, :] = s
        Sx = 1.0 + 1j * sigma_x / (self.epsilon0 * self.omega)
        for pre in ("eps", "mu"):
            xx, yy, zz = (self._cell_material_array(pre, c) for c in _COMPS)
            xx *= 1 / Sx
            yy *= Sx
            zz *= Sx
        self.update_component_materials()
        self._invalidate_solution()

    # ---- cell -> field averaging (Periodic_Solver_2D.py:375-482) ---------------------------------------------
    @staticmethod
    def _flat(values):
        return values.ravel(order="F")

    @staticmethod
    def _diag(values):
        return sp.diags(values.ravel(order="F"), format="csr")

    def _average_periodic_z(self, values, keep=None):
        out = 0.5 * (values + np.roll(values, -1, axis=1))
        if keep is not None:
            ii, jj = np.nonzero(keep)
            out[ii, jj] = values[ii, jj]
            out[ii, (jj - 1) % self.Nz] = values[ii, jj]
        return out

    def _average_x(self, values, keep=None):
        acc = np.zeros((self.Nx + 1, self.Nz), dtype=complex)
        cnt = np.zeros((self.Nx + 1, self.Nz), dtype=float)
        for o in (0, 1):
            acc[o:o + self.Nx, :] += values
            cnt[o:o + self.Nx, :] += 1
        acc = acc / cnt
        if keep is not None:
            ii, jj = np.nonzero(keep)
            for o in (0, 1):
                acc[ii + o, jj] = values[ii, jj]
        return acc

    def _material_on_fields(self, eps_r_xx, eps_r_yy, eps_r_zz, mu_r_xx, mu_r_yy, mu_r_zz, no_average_mask):
        z, x = self._average_periodic_z, self._average_x
        return {"eps_xx": z(eps_r_xx, no_average_mask), "eps_yy": x(eps_r_yy, no_average_mask),
                "eps_zz": x(eps_r_zz, no_average_mask), "mu_xx": x(mu_r_xx, no_average_mask),
                "mu_yy": z(mu_r_yy, no_average_mask), "mu_zz": mu_r_zz.copy()}

    def _set_component_materials(self, materials):
        for key, val in materials.items():
            pre, comp = key.split("_")
            setattr(self, f"{pre}_r_{comp}", val.copy())

    def _cells(self, copy=False):
        arrs = [self._cell_material_array(pre, comp) for pre in ("eps", "mu") for comp in _COMPS]
        return [a.copy() for a in arrs] if copy else arrs

    def update_component_materials(self):
        materials = self._material_on_fields(*self._cells(), self.material_no_average_mask)
        self._set_component_materials(materials)
        return materials

    def _effective_materials_and_masks(self):
        """Non-finite material entries become PEC / PMC masks, constrained entries are set to 1 (:436-482)."""
        materials = self._material_on_fields(*self._cells(copy=True), self.material_no_average_mask.copy())
        out, masks = [], []
        for pre, kind in (("eps", "pec"), ("mu", "pmc")):
            for comp in _COMPS:
                vals = materials[f"{pre}_{comp}"]
                mask = self._component_array(kind, comp) | ~np.isfinite(vals)
                vals[mask] = 1.0 + 0j
                out.append(vals)
                masks.append(mask)
        self._set_component_materials(materials)
        return (*out, *masks)

    # ---- pencils (Periodic_Solver_2D.py:484-521) --------------------------------------------------------------
    def _build_tm_system(self, eps_r_xx, eps_r_zz, mu_r_yy):
        w = self.omega
        zero = sp.csr_matrix((self.n_ex, self.n_hy), dtype=complex)
        D1 = 1j * w * self.mu0 * self._diag(mu_r_yy)
        D1 += 1j / w * self.DEX_EZ_TO_EX @ (1 / self.epsilon0 * self._diag(eps_r_zz).power(-1) @ self.DHX_HY_TO_EZ)
        D2 = 1j * w * self.epsilon0 * self._diag(eps_r_xx)
        return (sp.bmat([[self.DEZ_EX, D1], [D2, self.DHZ_HY]], format="csr"),
                sp.bmat([[self.AZ_EX, zero], [zero, self.AZ_HY]], format="csr"))

    def _build_te_system(self, eps_r_yy, mu_r_xx, mu_r_zz):
        w = self.omega
        zero = sp.csr_matrix((self.n_hx, self.n_ey), dtype=complex)
        D1 = -1j * w * self.epsilon0 * self._diag(eps_r_yy)
        D1 -= 1j / w * self.DHX_HZ_TO_HX @ (1 / self.mu0 * self._diag(mu_r_zz).power(-1) @ self.DEX_EY_TO_HZ)
        D2 = -1j * w * self.mu0 * self._diag(mu_r_xx)
        return (sp.bmat([[self.DHZ_HX, D1], [D2, self.DEZ_EY]], format="csr"),
                sp.bmat([[self.AZ_HX, zero], [zero, self.AZ_EY]], format="csr"))

    def _free_mask(self, pec_xx_mask, pec_yy_mask, pmc_xx_mask, pmc_yy_mask):
        if self.polarization == "TM":
            return np.concatenate((~self._flat(pec_xx_mask), ~self._flat(pmc_yy_mask)))
        return np.concatenate((~self._flat(pmc_xx_mask), ~self._flat(pec_yy_mask)))

    def _pencil(self):
        """(A, B, free) exactly as `solve` of the reference forms them (:529-554)."""
        (exx, eyy, ezz, mxx, myy, mzz, pxx, pyy, _pzz, qxx, qyy, _qzz) = self._effective_materials_and_masks()
        A, B = self._build_tm_system(exx, ezz, myy) if self.polarization == "TM" else self._build_te_system(eyy, mxx, mzz)
        free = self._free_mask(pxx, pyy, qxx, qyy)
        if np.count_nonzero(free) <= self.num_modes:
            raise ValueError(f"Not enough unconstrained DOFs to solve {self.num_modes} modes.")
        return A[free, :][:, free], B[free, :][:, free], free

    def _unknown_planes(self):
        """z index of every unknown of the two stacked fields (flat index i + nx j)."""
        nx = self.Nx if self.polarization == "TM" else self.Nx + 1
        one = np.arange(nx * self.Nz) // nx
        return np.concatenate((one, one))

    # ---- solve (Periodic_Solver_2D.py:523-588) ----------------------------------------------------------------
    def solve(self, guess=None, tol=None, ncv=None):
        import torch
        from .operators import CsrOperator

        guess = self.guess if guess is None else guess
        tol = self.tol if tol is None else tol
        ncv = self.ncv if ncv is None else ncv
        A, B, free = self._pencil()
        dev = torch.device(self.device or "cuda")
        # Equilibrate the pencil before the eigen-solve (eigenvalues are invariant under A -> Dr A Dc, B -> Dr B Dc).
        # With H in A/m, E in V/m and 1e8 penalty materials the entries of A span 1e-6 ... 1e13: a Ritz pair converged
        # to 1e-12 in the shift-invert operator is then only a 1e-7 eigenpair of the pencil, and a dense LAPACK
        # eigen-solve reaches no better than 1e-9 either; after Ruiz scaling both reach 1e-12 ... 1e-14.
        dr, dc = _eigs.ruiz_equilibrate(abs(A - guess * B) + abs(B))
        As = (sp.diags(dr) @ A @ sp.diags(dc)).tocsr()
        Bs = (sp.diags(dr) @ B @ sp.diags(dc)).tocsr()
        planes = self._unknown_planes()[free] if self.PLANE_BLOCK else None
        Ad, Bd = CsrOperator(As, device=dev), CsrOperator(Bs, device=dev)
        shifted = _eigs.make_shifted_solver(As, Bs, guess, dev, dense_limit=self.DENSE_LIMIT, planes=planes, nplanes=self.Nz)
        v0 = torch.from_numpy((1.0 / dc).astype(complex)).to(dev)          # the reference's start vector (ones), balanced
        lam, X, res, restarts = _eigs.eigs_shift_invert(Ad.apply, Bd.apply, shifted, A.shape[0], self.num_modes, guess,
                                                         ncv=ncv, tol=max(float(tol or 0.0), 1e-12), v0=v0, device=dev)
        order = np.argsort(np.abs(lam - guess))
        lam = lam[order]
        # relative residual of every pair in the reference's (unscaled) pencil, on the device:
        # A x = Dr^-1 (A' x'), B x = Dr^-1 (B' x') for x = Dc x'
        inv_dr = torch.from_numpy(1.0 / dr).to(dev)
        resid = np.zeros(len(lam))
        for pos, i in enumerate(order):
            ax, bx = Ad.apply(X[i]) * inv_dr, Bd.apply(X[i]) * inv_dr
            den = float(torch.linalg.vector_norm(ax)) + abs(lam[pos]) * float(torch.linalg.vector_norm(bx))
            resid[pos] = float(torch.linalg.vector_norm(ax - complex(lam[pos]) * bx)) / den if den > 0 else 0.0
        X = X.T.cpu().numpy()[:, order] * dc[:, None]                       # back to the reference's variables
        X /= np.linalg.norm(X, axis=0, keepdims=True)
        self.solver_info = dict(residuals=resid, balanced_residuals=res[order], restarts=restarts, n=int(A.shape[0]), sigma=guess)
        eigenvectors = np.zeros((free.size, self.num_modes), dtype=complex)
        eigenvectors[free, :] = X
        self.eigenvalues = lam
        self.eigenvectors = eigenvectors
        self.neff = self.eigenvalues / self.k0
        self.propagation_constant = np.imag(self.neff)
        self.attenuation_constant = np.real(self.neff)
        first, second = ("Ex", "Hy") if self.polarization == "TM" else ("Hx", "Ey")
        n_first = self.n_ex if self.polarization == "TM" else self.n_hx
        setattr(self, first, eigenvectors[:n_first, :])
        setattr(self, second, eigenvectors[n_first:, :])
        self.spurious_scores = np.zeros(self.num_modes, dtype=float)
        self.accepted_candidate_indices = np.arange(self.num_modes)
        self.rejected_candidate_indices = np.array([], dtype=int)
        self.unselected_candidate_indices = np.array([], dtype=int)

    def _field_array(self, field, mode, shape):
        return field[:, mode].reshape(shape, order="F")

    def visualize_with_gui(self):
        raise NotImplementedError("the Tk / matplotlib viewer of the reference is outside the GPU hot path")
