"""GPU drop-in for the reference's ``ModeSolver2D`` (Mode_Solver_2D/Mode_Solver_2D.py:12-1036):
full-vector waveguide modes on a true staggered Yee grid.  Same constructor, geometry / boundary
helpers (`add_rectangle`, `add_circle`, `add_triangle`, `add_pec`, `add_pmc`, `add_pml`/`add_UPML`,
`add_impedance_surface`), same `solve(sigma=None)` results (`eigenvalues`, `eigenvectors`, `neff`,
`propagation_constant`, `attenuation_constant`, `Ex..Hz` with shapes `(*grid, num_modes)`).

Hot path on the GPU:
  * the eigenproblem ``Omega e = lambda e`` with ``Omega = P_red Q_red`` (Mode_Solver_2D.py:752-774)
    keeps ``P_red`` and ``Q_red`` as two device CSR matrices — the product is never formed — and
    ``eigs(Omega, k, sigma)`` (ARPACK + SuperLU, :780) becomes the Krylov-Schur shift-invert solver
    of ``fdfd_b200/eigs.py`` with ``(Omega - sigma I)^-1`` applied by Jacobi-preconditioned
    BiCGSTAB (csrc/krylov.cu) on the factored operator (csrc/csr_ops.cu).
Host side (numpy, O(N) setup exactly as the reference): sub-pixel geometry, UPML scaling,
cell -> Yee averaging, PEC/PMC masks, assembly of the sparse factors, field reconstruction.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import _lib
from . import eigs as _eigs
from .operators import CsrOperator, CsrProductOperator

_COMPS = ("xx", "yy", "zz")


class ModeSolver2D:
    epsilon0 = 8.854187817e-12           # Mode_Solver_2D.py:26-29
    mu0 = 4e-7 * np.pi

    def __init__(self, frequency, x_range, y_range, Nx, Ny, num_modes, guess=None, *, device=None):
        self.frequency, self.x_range, self.y_range = frequency, x_range, y_range
        self.Nx, self.Ny = int(Nx), int(Ny)
        if self.Nx <= 0 or self.Ny <= 0:
            raise ValueError("Nx and Ny must be positive.")
        self.dx, self.dy = x_range / self.Nx, y_range / self.Ny
        self.c = 1 / np.sqrt(self.epsilon0 * self.mu0)
        self.k_0 = 2 * np.pi * frequency / self.c
        self.dx_normalized, self.dy_normalized = self.k_0 * self.dx, self.k_0 * self.dy
        nx, ny = self.Nx, self.Ny
        self.shape_cell = (nx, ny)
        self.shape_ex = self.shape_hy = (nx, ny + 1)
        self.shape_ey = self.shape_hx = (nx + 1, ny)
        self.shape_ez = (nx + 1, ny + 1)
        self.shape_hz = (nx, ny)
        self.n_ex = self.n_hy = nx * (ny + 1)
        self.n_ey = self.n_hx = (nx + 1) * ny
        self.n_ez = (nx + 1) * (ny + 1)
        self.n_hz = nx * ny
        self.n_e = self.n_ex + self.n_ey
        self.n_h = self.n_hx + self.n_hy
        for pre in ("eps", "mu"):
            for c in _COMPS:
                setattr(self, f"cell_{pre}_r_{c}", np.ones(self.shape_cell, dtype=complex))
        self.material_no_average_mask = np.zeros(self.shape_cell, dtype=bool)
        e_shapes = dict(xx=self.shape_ex, yy=self.shape_ey, zz=self.shape_ez)
        h_shapes = dict(xx=self.shape_hx, yy=self.shape_hy, zz=self.shape_hz)
        for c in _COMPS:
            setattr(self, f"eps_r_{c}", np.ones(e_shapes[c], dtype=complex))
            setattr(self, f"mu_r_{c}", np.ones(h_shapes[c], dtype=complex))
            setattr(self, f"pec_{c}_mask", np.zeros(e_shapes[c], dtype=bool))
            setattr(self, f"pmc_{c}_mask", np.zeros(h_shapes[c], dtype=bool))
        self._pec_regions, self._pmc_regions = [], []
        self.num_modes = int(num_modes)
        if self.num_modes <= 0:
            raise ValueError("num_modes must be positive.")
        self.guess = guess
        self._auto_guess = guess is None
        if self._auto_guess:
            self.guess = self._default_guess()
        self.device = device
        self.solver_info = {}
        self._invalidate_solution()

    # ---- small helpers ------------------------------------------------------------------------------
    def _cells(self, pre):
        return [getattr(self, f"cell_{pre}_r_{c}") for c in _COMPS]

    def _default_guess(self):
        """-max |eps, mu| over the finite cell values (Mode_Solver_2D.py:84-98)."""
        best = 0.0
        for arr in self._cells("eps") + self._cells("mu"):
            mag = np.abs(arr)
            mag = mag[np.isfinite(mag)]
            if mag.size:
                best = max(best, np.max(mag))
        return -best

    def _resolve_eigs_guess(self, sigma):
        if sigma is not None:
            return sigma
        if self._auto_guess:
            self.guess = self._default_guess()
        return self.guess

    def _invalidate_solution(self):
        self.eigenvalues = self.eigenvectors = self.neff = None
        self.propagation_constant = self.attenuation_constant = None
        for name in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"):
            self.__dict__.pop(name, None)

    @staticmethod
    def _three(name, value):
        if np.isscalar(value):
            return np.full(3, value, dtype=complex)
        arr = np.asarray(value, dtype=complex)
        if arr.ndim == 1 and arr.size == 3:
            return arr
        raise ValueError(f"{name} must be a scalar or a length-3 1D array (xx, yy, zz).")

    @staticmethod
    def _validate_components(components):
        comps = (components,) if isinstance(components, str) else tuple(components)
        bad = set(comps) - set(_COMPS)
        if bad:
            raise ValueError(f"components contains invalid tensor component(s): {sorted(bad)}.")
        return comps

    def _step(self, axis):
        return self.dx if axis == "x" else self.dy

    def _to_index(self, value, axis):
        if isinstance(value, (int, np.integer)):
            return int(value)
        if isinstance(value, (float, np.floating)):
            return int(round(float(value) / self._step(axis)))
        raise ValueError("Region bounds must be int grid indices or float physical positions in metres.")

    def _to_length(self, value, axis):
        if isinstance(value, (int, np.integer)):
            return int(value) * self._step(axis)
        if isinstance(value, (float, np.floating)):
            return float(value)
        raise ValueError("Coordinates must be int grid indices or float physical positions in metres.")

    def _point(self, name, point):
        try:
            x, y = point
        except (TypeError, ValueError):
            raise ValueError(f"{name} must be an (x, y) pair.")
        return self._to_length(x, "x"), self._to_length(y, "y")

    def _span(self, name, values, axis):
        try:
            lo, hi = values
        except (TypeError, ValueError):
            raise ValueError(f"{name} must be a (min, max) pair.")
        lo, hi = self._to_length(lo, axis), self._to_length(hi, axis)
        if hi <= lo:
            raise ValueError(f"{name} must satisfy max > min.")
        if lo < 0 or hi > (self.x_range if axis == "x" else self.y_range):
            raise ValueError(f"{name} is out of bounds of the simulation grid.")
        return lo, hi

    def _radius(self, name, r):
        if isinstance(r, (int, np.integer)):
            return int(r) * min(self.dx, self.dy)
        if isinstance(r, (float, np.floating)):
            return float(r)
        raise ValueError(f"{name} must be an int number of cells or a float physical radius in metres.")

    def _region_slices(self, x_range, y_range):
        try:
            x0, x1 = self._to_index(x_range[0], "x"), self._to_index(x_range[1], "x")
            y0, y1 = self._to_index(y_range[0], "y"), self._to_index(y_range[1], "y")
        except (TypeError, IndexError):
            raise ValueError("x_range and y_range must be (min, max) pairs.")
        if not (x1 > x0 and y1 > y0):
            raise ValueError("x_range and y_range must satisfy max > min.")
        if not (0 <= x0 < x1 <= self.Nx and 0 <= y0 < y1 <= self.Ny):
            raise ValueError("Region is out of bounds of the simulation grid.")
        return slice(x0, x1), slice(y0, y1)

    def _bbox(self, xmin, xmax, ymin, ymax):
        return (max(0, int(np.floor(xmin / self.dx))), min(self.Nx, int(np.ceil(xmax / self.dx))),
                max(0, int(np.floor(ymin / self.dy))), min(self.Ny, int(np.ceil(ymax / self.dy))))

    @staticmethod
    def _subpixel_axis(start, stop, step, sub):
        idx = np.arange(start, stop, dtype=float)
        off = (np.arange(sub, dtype=float) + 0.5) / sub
        return (idx[:, None] + off[None, :]) * step

    # ---- geometry (Mode_Solver_2D.py:225-339) -----------------------------------------------------------
    def _blend(self, epsilon, mu, fraction, sl_x, sl_y):
        epsilon, mu = self._three("epsilon", epsilon), self._three("mu", mu)
        fraction = np.asarray(fraction, dtype=float)
        if fraction.shape != self.cell_eps_r_xx[sl_x, sl_y].shape:
            raise ValueError("fraction shape does not match target cell region.")
        hit = fraction > 0.0
        if not np.any(hit):
            return
        for pre, vals in (("eps", epsilon), ("mu", mu)):
            for arr, v in zip(self._cells(pre), vals):
                tgt = arr[sl_x, sl_y]
                tgt[hit] = tgt[hit] * (1.0 - fraction[hit]) + v * fraction[hit]
        self.material_no_average_mask[sl_x, sl_y][hit] = False
        self.update_component_materials()
        self._invalidate_solution()

    def _fill(self, epsilon, mu, bbox, inside, subpixels):
        x0, x1, y0, y1 = bbox
        if x0 >= x1 or y0 >= y1:
            return
        xs = self._subpixel_axis(x0, x1, self.dx, subpixels)
        ys = self._subpixel_axis(y0, y1, self.dy, subpixels)
        self._blend(epsilon, mu, inside(xs, ys).mean(axis=(1, 3)), slice(x0, x1), slice(y0, y1))

    @staticmethod
    def _subpixels(n):
        n = int(n)
        if n <= 0:
            raise ValueError("subpixels must be positive.")
        return n

    def add_rectangle(self, epsilon, mu, x_range, y_range, *, subpixels=8):
        xmin, xmax = self._span("x_range", x_range, "x")
        ymin, ymax = self._span("y_range", y_range, "y")
        subpixels = self._subpixels(subpixels)

        def inside(xs, ys):
            return ((xs >= xmin) & (xs <= xmax))[:, :, None, None] & ((ys >= ymin) & (ys <= ymax))[None, None, :, :]
        self._fill(epsilon, mu, self._bbox(xmin, xmax, ymin, ymax), inside, subpixels)

    def add_circle(self, epsilon, mu, center, r1, r2=None, *, subpixels=8):
        cx, cy = self._point("center", center)
        r1 = self._radius("r1", r1)
        r2 = 0.0 if r2 is None else self._radius("r2", r2)
        subpixels = self._subpixels(subpixels)
        if r1 <= 0:
            raise ValueError("r1 must be positive.")
        if r2 < 0 or r2 >= r1:
            raise ValueError("r2 must be non-negative and smaller than r1.")

        def inside(xs, ys):
            rr = (xs[:, :, None, None] - cx) ** 2 + (ys[None, None, :, :] - cy) ** 2
            m = rr <= r1 ** 2
            if r2 > 0:
                m &= rr >= r2 ** 2
            return m
        self._fill(epsilon, mu, self._bbox(cx - r1, cx + r1, cy - r1, cy + r1), inside, subpixels)

    def add_triangle(self, epsilon, mu, p1, p2, p3, *, subpixels=8):
        p1, p2, p3 = (self._point(n, p) for n, p in (("p1", p1), ("p2", p2), ("p3", p3)))
        subpixels = self._subpixels(subpixels)
        area2 = (p2[0] - p1[0]) * (p3[1] - p1[1]) - (p3[0] - p1[0]) * (p2[1] - p1[1])
        if abs(area2) <= 1e-30:
            raise ValueError("Triangle points must not be collinear.")

        def inside(xs, ys):
            x, y = xs[:, :, None, None], ys[None, None, :, :]
            d = [(x - b[0]) * (a[1] - b[1]) - (a[0] - b[0]) * (y - b[1]) for a, b in ((p1, p2), (p2, p3), (p3, p1))]
            neg = (d[0] < 0) | (d[1] < 0) | (d[2] < 0)
            pos = (d[0] > 0) | (d[1] > 0) | (d[2] > 0)
            return ~(neg & pos)
        xs_, ys_ = (p1[0], p2[0], p3[0]), (p1[1], p2[1], p3[1])
        self._fill(epsilon, mu, self._bbox(min(xs_), max(xs_), min(ys_), max(ys_)), inside, subpixels)

    # ---- PEC / PMC (Mode_Solver_2D.py:341-402) -----------------------------------------------------------
    def component_masks_from_cell_mask(self, cell_mask, field="electric"):
        mask = np.asarray(cell_mask, dtype=bool)
        if mask.shape != self.shape_cell:
            raise ValueError(f"cell_mask must have shape {self.shape_cell}.")
        if field not in ("electric", "magnetic"):
            raise ValueError("field must be 'electric' or 'magnetic'.")
        ii, jj = np.nonzero(mask)
        if field == "electric":
            xx, yy, zz = (np.zeros(s, bool) for s in (self.shape_ex, self.shape_ey, self.shape_ez))
            for o in (0, 1):
                xx[ii, jj + o] = True
                yy[ii + o, jj] = True
                zz[ii + o, jj] = True
                zz[ii + o, jj + 1] = True
        else:
            xx, yy, zz = (np.zeros(s, bool) for s in (self.shape_hx, self.shape_hy, self.shape_hz))
            for o in (0, 1):
                xx[ii + o, jj] = True
                yy[ii, jj + o] = True
            zz[ii, jj] = True
        return xx, yy, zz

    def _add_wall(self, kind, x_range, y_range, components):
        sl = self._region_slices(x_range, y_range)
        (self._pec_regions if kind == "pec" else self._pmc_regions).append(sl)
        cells = np.zeros(self.shape_cell, dtype=bool)
        cells[sl] = True
        masks = self.component_masks_from_cell_mask(cells, field="electric" if kind == "pec" else "magnetic")
        chosen = _COMPS if components is None else self._validate_components(components)
        for c, m in zip(_COMPS, masks):
            if c in chosen:
                getattr(self, f"{kind}_{c}_mask")[:] |= m
        self._effective_materials_and_masks()
        self._invalidate_solution()

    def add_pec(self, x_range, y_range, components=None):
        self._add_wall("pec", x_range, y_range, components)

    def add_pmc(self, x_range, y_range, components=None):
        self._add_wall("pmc", x_range, y_range, components)

    # ---- absorbing layers / sheets (Mode_Solver_2D.py:404-476) -------------------------------------------
    def add_pml(self, pml_width=50, n=3, sigma_max=5, direction="all"):
        pml_width = int(pml_width)
        if pml_width <= 0:
            raise ValueError("pml_width must be positive.")
        if direction not in ("x-", "x+", "x", "y-", "y+", "y", "all"):
            raise ValueError("direction must be one of 'x-', 'x+', 'x', 'y-', 'y+', 'y', or 'all'.")
        sig = {"x": np.zeros(self.shape_cell), "y": np.zeros(self.shape_cell)}
        prof = [sigma_max * ((pml_width - i) / pml_width) ** n for i in range(pml_width)]
        for ax, N in (("x", self.Nx), ("y", self.Ny)):
            for side in ("-", "+"):
                if direction not in (ax + side, ax, "all"):
                    continue
                for i in range(min(pml_width, N)):
                    k = i if side == "-" else -i - 1
                    if ax == "x":
                        sig["x"][k, :] = prof[i]
                    else:
                        sig["y"][:, k] = prof[i]
        omega = 2 * np.pi * self.frequency
        Sx = 1.0 + 1j * sig["x"] / (self.epsilon0 * omega)
        Sy = 1.0 + 1j * sig["y"] / (self.epsilon0 * omega)
        for pre in ("eps", "mu"):
            xx, yy, zz = self._cells(pre)
            xx *= Sy / Sx
            yy *= Sx / Sy
            zz *= Sx * Sy
        self.update_component_materials()
        self._invalidate_solution()

    def add_UPML(self, pml_width=50, n=3, sigma_max=5, direction="all"):
        self.add_pml(pml_width=pml_width, n=n, sigma_max=sigma_max, direction=direction)

    def add_impedance_surface(self, Zs, position, *, orientation="x", thickness_cells=1,
                              eps_components=("xx", "yy", "zz")):
        if orientation not in ("x", "y"):
            raise ValueError("orientation must be 'x' or 'y'.")
        thickness_cells = int(thickness_cells)
        if thickness_cells <= 0:
            raise ValueError("thickness_cells must be positive.")
        k = self._to_index(position, orientation)
        if orientation == "x":
            xr, yr, thick = (k, k + thickness_cells), (0, self.Ny), thickness_cells * self.dx
        else:
            xr, yr, thick = (0, self.Nx), (k, k + thickness_cells), thickness_cells * self.dy
        sl = self._region_slices(xr, yr)
        delta = -1j / (2 * np.pi * self.frequency * self.epsilon0 * thick * Zs)
        for c in self._validate_components(eps_components):
            getattr(self, f"cell_eps_r_{c}")[sl] += delta
        self.update_component_materials()
        self._invalidate_solution()

    # ---- cell -> Yee averaging (Mode_Solver_2D.py:546-624) ------------------------------------------------
    def _average(self, values, shape, offsets, keep):
        acc = np.zeros(shape, dtype=complex)
        cnt = np.zeros(shape, dtype=float)
        nx, ny = self.shape_cell
        for ox, oy in offsets:
            acc[ox:ox + nx, oy:oy + ny] += values
            cnt[ox:ox + nx, oy:oy + ny] += 1
        acc = acc / cnt
        if keep is not None:
            ii, jj = np.nonzero(keep)
            for ox, oy in offsets:
                acc[ii + ox, jj + oy] = values[ii, jj]
        return acc

    _ALONG_Y = ((0, 0), (0, 1))
    _ALONG_X = ((0, 0), (1, 0))
    _NODE = ((0, 0), (1, 0), (0, 1), (1, 1))

    def _material_on_fields(self, exx, eyy, ezz, mxx, myy, mzz, keep):
        return {"eps_xx": self._average(exx, self.shape_ex, self._ALONG_Y, keep),
                "eps_yy": self._average(eyy, self.shape_ey, self._ALONG_X, keep),
                "eps_zz": self._average(ezz, self.shape_ez, self._NODE, keep),
                "mu_xx": self._average(mxx, self.shape_hx, self._ALONG_X, keep),
                "mu_yy": self._average(myy, self.shape_hy, self._ALONG_Y, keep),
                "mu_zz": mzz.copy()}

    def _store(self, mats):
        for pre in ("eps", "mu"):
            for c in _COMPS:
                setattr(self, f"{pre}_r_{c}", mats[f"{pre}_{c}"].copy())

    def update_component_materials(self):
        mats = self._material_on_fields(*self._cells("eps"), *self._cells("mu"), self.material_no_average_mask)
        self._store(mats)
        return mats

    def _effective_materials_and_masks(self):
        """Mode_Solver_2D.py:626-676."""
        cells = {pre: [a.copy() for a in self._cells(pre)] for pre in ("eps", "mu")}
        pec = [getattr(self, f"pec_{c}_mask").copy() for c in _COMPS]
        pmc = [getattr(self, f"pmc_{c}_mask").copy() for c in _COMPS]
        for pre, walls, fld in (("eps", pec, "electric"), ("mu", pmc, "magnetic")):
            for k, vals in enumerate(cells[pre]):
                bad = ~np.isfinite(vals)
                if np.any(bad):
                    walls[k][:] |= self.component_masks_from_cell_mask(bad, field=fld)[k]
                    vals[bad] = 1.0 + 0j
        mats = self._material_on_fields(*cells["eps"], *cells["mu"], self.material_no_average_mask.copy())
        for k, c in enumerate(_COMPS):
            mats[f"eps_{c}"][pec[k]] = 1.0 + 0j
            mats[f"mu_{c}"][pmc[k]] = 1.0 + 0j
        self._store(mats)
        return (mats, *pec, *pmc)

    # ---- staggered difference operators (Mode_Solver_2D.py:482-544) ---------------------------------------
    @staticmethod
    def _difference(in_shape, out_shape, scale, axis, with_h=True):
        """Rectangular forward difference between two staggered grids and its H-field counterpart -D^H,
        assembled by the K1b kernel (csrc/yee_assemble.cu: stagger_fill_kernel); byte-identical to the
        reference's Python double loop + coo -> csr (Mode_Solver_2D.py:482-514)."""
        from . import stagger
        return stagger.difference(in_shape, out_shape, scale, axis, with_h=with_h)

    def _yeeder2d(self):
        dx, dy = self.dx_normalized, self.dy_normalized
        self.DEX_EZ_TO_EX, self.DHX_HY_TO_EZ = self._difference(self.shape_ez, self.shape_ex, dx, 0)
        self.DEY_EZ_TO_EY, self.DHY_HX_TO_EZ = self._difference(self.shape_ez, self.shape_ey, dy, 1)
        self.DEX_EY_TO_HZ, self.DHX_HZ_TO_HX = self._difference(self.shape_ey, self.shape_hz, dx, 0)
        self.DEY_EX_TO_HZ, self.DHY_HZ_TO_HY = self._difference(self.shape_ex, self.shape_hz, dy, 1)
        self.DEX, self.DEY, self.DHX, self.DHY = (self.DEX_EY_TO_HZ, self.DEY_EX_TO_HZ, self.DHX_HY_TO_EZ,
                                                  self.DHY_HX_TO_EZ)
        return (self.DEX_EZ_TO_EX, self.DEY_EZ_TO_EY, self.DEX_EY_TO_HZ, self.DEY_EX_TO_HZ,
                self.DHX_HY_TO_EZ, self.DHY_HX_TO_EZ, self.DHX_HZ_TO_HX, self.DHY_HZ_TO_HY)

    # ---- solve (Mode_Solver_2D.py:727-814) ----------------------------------------------------------------
    def _has_lossy_material(self):
        for vals in self._cells("eps") + self._cells("mu"):
            fin = np.isfinite(vals)
            if np.any(np.abs(np.imag(vals[fin])) > 1e-14):
                return True
        return False

    def _passive_positive_neff(self, neff_squared):
        root = np.sqrt(neff_squared)
        neff = np.where(np.real(root) < 0, -root, root)
        re, im = np.real(neff), np.imag(neff)
        tol = 1e-12 * np.maximum(1.0, np.abs(neff))
        re = np.where(np.abs(re) <= tol, 0.0, re)
        im = np.where(np.abs(im) <= tol, 0.0, np.abs(im))
        if not self._has_lossy_material():
            im = np.zeros_like(im)
        return re + 1j * im

    def _factors(self):
        """P_red, Q_red and the bookkeeping the reconstruction needs (host, sparse)."""
        mats, pxx, pyy, pzz, mxx, myy, mzz = self._effective_materials_and_masks()
        F = lambda a: a.ravel(order="F")  # noqa: E731
        dg = lambda a: sp.diags(F(a), format="csr")  # noqa: E731

        def inv_free(vals, fixed):
            flat, fx = F(vals), F(fixed)
            inv = np.zeros_like(flat, dtype=complex)
            inv[~fx] = 1.0 / flat[~fx]
            return sp.diags(inv, format="csr")
        ezi, mzi = inv_free(mats["eps_zz"], pzz), inv_free(mats["mu_zz"], mzz)
        d_ez_ex, d_ez_ey, d_ey_hz, d_ex_hz, d_hy_ez, d_hx_ez, d_hz_hx, d_hz_hy = self._yeeder2d()
        P = sp.bmat([[d_ez_ex @ ezi @ d_hx_ez, -(d_ez_ex @ ezi @ d_hy_ez + dg(mats["mu_yy"]))],
                     [d_ez_ey @ ezi @ d_hx_ez + dg(mats["mu_xx"]), -d_ez_ey @ ezi @ d_hy_ez]], format="csr")
        Q = sp.bmat([[d_hz_hx @ mzi @ d_ex_hz, -(d_hz_hx @ mzi @ d_ey_hz + dg(mats["eps_yy"]))],
                     [d_hz_hy @ mzi @ d_ex_hz + dg(mats["eps_xx"]), -d_hz_hy @ mzi @ d_ey_hz]], format="csr")
        free_e = np.concatenate((~F(pxx), ~F(pyy)))
        free_h = np.concatenate((~F(mxx), ~F(myy)))
        Q_red = Q[free_h, :]
        return dict(P_e=P[:, free_h][free_e, :], Q_e=Q_red[:, free_e], Q_red=Q_red, free_e=free_e, free_h=free_h,
                    ezi=ezi, mzi=mzi, ops=(d_ey_hz, d_ex_hz, d_hy_ez, d_hx_ez), masks=(pxx, pyy, pzz, mxx, myy, mzz))

    def solve(self, sigma=None, *, tol=1e-11, inner_rtol=1e-12):
        import torch

        sigma = self._resolve_eigs_guess(sigma)
        f = self._factors()
        n_free = int(f["free_e"].sum())
        if n_free <= self.num_modes:
            raise ValueError(f"Not enough unconstrained electric DOFs ({n_free}) to solve {self.num_modes} modes.")
        # Omega = P_e Q_e on the free electric DOFs, kept factored on the device
        Om = CsrProductOperator(f["P_e"], f["Q_e"], shift=0.0, device=self.device)
        Oms = CsrProductOperator(f["P_e"], f["Q_e"], shift=-complex(sigma), device=self.device, diag=Om.diag)
        stats = dict(inner_solves=0, inner_iterations=0)

        def shifted(r):
            # the inner tolerance sits close to the rounding floor of b - (Omega - sigma) x: a solve that stalls
            # within two digits of it is as accurate as this operator allows and is accepted (recorded below)
            x, info = Oms.solve(r, method="bicgstab", rtol=inner_rtol, maxit=50000, raise_on_noconv=False)
            if not info["converged"] and not info["relres"] <= 100 * inner_rtol:
                raise _lib.NoConvergence(f"inner BiCGSTAB stalled at relres {info['relres']:.3e}", info)
            stats["inner_solves"] += 1
            stats["inner_iterations"] += info["iterations"]
            stats["inner_max_relres"] = max(stats.get("inner_max_relres", 0.0), info["relres"])
            return x
        lam, X, res, restarts = _eigs.eigs_shift_invert(Om.apply, lambda v: v, shifted, n_free, self.num_modes,
                                                         complex(sigma), tol=tol, device=Om.device,
                                                         dtype=torch.complex128)
        self.solver_info = dict(restarts=restarts, residuals=res, n=n_free, sigma=sigma, **stats)
        eigenvalues = lam
        order = np.argsort(np.real(eigenvalues))
        self.eigenvalues = eigenvalues[order]
        self.neff = self._passive_positive_neff(-self.eigenvalues)
        self.propagation_constant = np.real(self.neff)
        self.attenuation_constant = np.imag(self.neff)
        root = np.sqrt(self.eigenvalues)
        if np.any(np.abs(root) < 1e-300):
            raise ValueError("Encountered a near-zero eigenvalue while reconstructing magnetic fields.")
        # field reconstruction (Mode_Solver_2D.py:792-805) on the device: H = Q E / sqrt(lambda) and the two curls
        # are sparse mat-vecs with the device CSR kernel on the Ritz vectors where the eigen-solver left them;
        # the six fields come back to the host once
        dev, k = Om.device, self.num_modes
        d_ey_hz, d_ex_hz, d_hy_ez, d_hx_ez = f["ops"]
        idx = lambda mask: torch.from_numpy(np.flatnonzero(mask)).to(dev)  # noqa: E731
        Exy_d = torch.zeros((k, self.n_e), dtype=torch.complex128, device=dev)
        Exy_d[:, idx(f["free_e"])] = X[torch.from_numpy(order.copy()).to(dev)]
        inv_root = torch.from_numpy(1.0 / root).to(dev).unsqueeze(1)
        Hxy_d = torch.zeros((k, self.n_h), dtype=torch.complex128, device=dev)
        Hxy_d[:, idx(f["free_h"])] = CsrOperator(f["Q_red"], device=dev).apply_rows(Exy_d) * inv_root
        ex_d, ey_d = Exy_d[:, :self.n_ex].contiguous(), Exy_d[:, self.n_ex:].contiguous()
        hx_d, hy_d = Hxy_d[:, :self.n_hx].contiguous(), Hxy_d[:, self.n_hx:].contiguous()
        dvec = lambda D: torch.from_numpy(np.ascontiguousarray(D.diagonal(), dtype=np.complex128)).to(dev)  # noqa: E731
        op = lambda D: CsrOperator(D, device=dev)  # noqa: E731
        ez_d = dvec(f["ezi"]) * (op(d_hy_ez).apply_rows(hy_d) - op(d_hx_ez).apply_rows(hx_d))
        hz_d = dvec(f["mzi"]) * (op(d_ey_hz).apply_rows(ey_d) - op(d_ex_hz).apply_rows(ex_d))
        host = lambda t: np.ascontiguousarray(t.cpu().numpy().T)  # noqa: E731  -> (n, k) as the reference stores them
        Exy, Hxy = host(Exy_d), host(Hxy_d)
        self.eigenvectors = Exy
        ex, ey = Exy[:self.n_ex, :], Exy[self.n_ex:, :]
        hx, hy = Hxy[:self.n_hx, :], Hxy[self.n_hx:, :]
        ez, hz = host(ez_d), host(hz_d)
        # NB: copies — in the reference `Ex`/`Ey` are *views* of `eigenvectors` (Mode_Solver_2D.py:797-798,
        # 807-808), so its most-real rotation (:711-725) hits them twice and leaves Ex, Ey and the
        # eigenvectors with a run-dependent phase; here all six fields share one consistent phase.
        unfl = lambda a, shape: np.array(a, dtype=complex).reshape((*shape, a.shape[1]), order="F")  # noqa: E731
        self.Ex, self.Ey, self.Ez = unfl(ex, self.shape_ex), unfl(ey, self.shape_ey), unfl(ez, self.shape_ez)
        self.Hx, self.Hy, self.Hz = unfl(hx, self.shape_hx), unfl(hy, self.shape_hy), unfl(hz, self.shape_hz)
        for fld, m in zip((self.Ex, self.Ey, self.Ez, self.Hx, self.Hy, self.Hz), f["masks"]):
            fld[m, :] = 0.0
        self._rotate_modes_to_most_real()

    @staticmethod
    def _most_real_phase(*fields):
        vals = np.concatenate([np.asarray(x).ravel() for x in fields])
        vals = vals[np.isfinite(vals)]
        if vals.size == 0 or np.max(np.abs(vals)) == 0:
            return 1.0 + 0j
        return np.exp(-0.5j * np.angle(np.sum(vals ** 2)))

    def _rotate_modes_to_most_real(self):
        names = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")
        for m in range(self.num_modes):
            ph = self._most_real_phase(*(getattr(self, n)[:, :, m] for n in names))
            for n in names:
                getattr(self, n)[:, :, m] *= ph
            self.eigenvectors[:, m] *= ph

    # ---- visualisation: the reference ships a Tk/matplotlib GUI (Mode_Solver_2D.py:842-1036); host side ----
    def visualize(self, field="Ex", mode=0, part="abs", ax=None, show=True):
        import matplotlib.pyplot as plt
        if self.eigenvalues is None:
            raise RuntimeError("Run solve() first")
        data = getattr(self, field)[:, :, mode]
        data = {"abs": np.abs, "real": np.real, "imag": np.imag}[part](data)
        if ax is None:
            _, ax = plt.subplots()
        im = ax.imshow(data.T, origin="lower", extent=[0, self.x_range, 0, self.y_range], aspect="auto")
        ax.set_title(f"{field} mode {mode}  n_eff={self.neff[mode]:.6f}")
        plt.colorbar(im, ax=ax)
        if show:
            plt.show()
        return ax

    def visualize_with_gui(self):
        raise NotImplementedError("the Tk viewer of the reference is not part of the GPU engine; use visualize()")
