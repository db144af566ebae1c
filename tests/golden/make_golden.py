"""Regenerates the golden fixtures in this directory from the UNMODIFIED reference
(/root/reference, imported read-only through oracle/ref_loader.py with matplotlib/tkinter
stubs).  Run in the build container only:  python tests/golden/make_golden.py

Fixtures (all produced by the reference's own code paths, scipy 1.18.1 / numpy 2.3.5):
  yee2d_kat.npz        yeeder2d outputs (indptr/indices/data of DEX, DEY, DHX, DHY) for the
                       edge-case grid of SURVEY.md §4.2 / §8c
  scatter_small.npz    FDFD2DScatteringSolver at 120x120 (eps_r=4 cylinder, UPML 20, mask 30):
                       A@x for a seeded x (TE and TM), right-hand sides and solved Ez / Hz
  scatter_cfg1.npz     BASELINE config 1 (example_scattering_by_cylinder.py, 300x300, TM) and the
                       same geometry solved TE: probe values, norms and a 10x-decimated field
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

YEE_CASES = [
    # (Nx, Ny, dx, dy, bcx, bcy, kx, ky)
    (4, 3, 0.5, 0.25, 0, 0, None, None),
    (4, 3, 0.5, 0.25, 1, 1, 0.3, 0.7),
    (7, 5, 0.13, 0.29, 1, 0, 1.1, -0.4),
    (7, 5, 0.13, 0.29, 0, 1, 1.1, -0.4),
    (1, 5, 0.2, 0.3, 0, 0, 0.9, 0.2),
    (5, 1, 0.2, 0.3, 1, 1, 0.9, 0.2),
    (1, 1, 0.2, 0.3, 0, 1, 0.5, 0.6),
    (2, 2, 1.0, 1.0, 1, 1, 0.0, 0.0),
    (16, 33, 0.05, 0.07, 1, 1, 2.5, 3.5),
]


def yee_fixture():
    yeeder2d = ref_loader.ref_yeeder2d("Scattering")
    out = {"cases": np.array([[np.nan if v is None else v for v in c] for c in YEE_CASES], dtype=float)}
    for n, (Nx, Ny, dx, dy, bx, by, kx, ky) in enumerate(YEE_CASES):
        kinc = None if kx is None else [kx, ky]
        ops = yeeder2d([Nx, Ny], [dx, dy], [bx, by], kinc)
        for name, m in zip(("DEX", "DEY", "DHX", "DHY"), ops):
            out[f"c{n}_{name}_indptr"] = m.indptr
            out[f"c{n}_{name}_indices"] = m.indices
            out[f"c{n}_{name}_data"] = m.data
            out[f"c{n}_{name}_format"] = np.array(m.format)
    np.savez_compressed(os.path.join(HERE, "yee2d_kat.npz"), **out)


def _ref_sim(N, eps, pw, mask, f0=10e9):
    mod = ref_loader.load("Scattering_Solver_2D")
    lam = 299792458 / f0
    L = N / 50 * lam
    sim = mod.FDFD2DScatteringSolver(frequency=f0, x_range=L, y_range=L, Nx=N, Ny=N)
    sim.add_object(er_tensor=eps, mr_tensor=1.0, region_mask=(sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=pw, sigma_max=10)
    sim.add_mask(value=mask)
    sim.add_source(src_type="plane_wave", angle_deg=45.0, polarization="TE")
    return sim


def scatter_small():
    N = 120
    sim = _ref_sim(N, 4.0, 20, 30)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N)
    out = dict(N=N, eps=4.0, pml=20, mask=30, x=x)
    for pol in ("TE", "TM"):
        F = getattr(sim, "solve_total_field_" + pol)()
        A = getattr(sim, "_A_" + pol)
        out[f"y_{pol}"] = A @ x
        out[f"b_{pol}"] = np.asarray((sim.Q @ A - A @ sim.Q) @ sim.source).ravel()
        out[f"F_{pol}"] = F
    np.savez_compressed(os.path.join(HERE, "scatter_small.npz"), **out)


def scatter_cfg1():
    N = 300
    sim = _ref_sim(N, -1e6, 40, 80)
    out = dict(N=N)
    for pol in ("TM", "TE"):
        F = getattr(sim, "solve_total_field_" + pol)()
        out[f"probe_{pol}"] = F[150, 75]
        out[f"max_{pol}"] = np.abs(F).max()
        out[f"norm_{pol}"] = np.linalg.norm(F)
        out[f"dec_{pol}"] = F[::10, ::10].copy()
    np.savez_compressed(os.path.join(HERE, "scatter_cfg1.npz"), **out)


def scatter_1024():
    """SURVEY.md §8c "scaled 1024^2" case (same script, L = N/50 lambda): probes, norm and a 32x-decimated
    field of the reference's direct (SuperLU) solve, TM and TE.  ~2 min, ~6 GB."""
    N = 1024
    sim = _ref_sim(N, -1e6, 40, 80)
    out = dict(N=N)
    for pol in ("TM", "TE"):
        F = getattr(sim, "solve_total_field_" + pol)()
        out[f"probe_{pol}"] = F[512, 256]
        out[f"max_{pol}"] = np.abs(F).max()
        out[f"norm_{pol}"] = np.linalg.norm(F)
        out[f"dec_{pol}"] = F[::32, ::32].copy()
    np.savez_compressed(os.path.join(HERE, "scatter_1024.npz"), **out)


def scatter_2048():
    """SURVEY.md §8c "scaled 2048^2" case: the largest size the reference's SuperLU solve fits in this
    container (17.6 GB, ~5 min per polarisation).  Besides probes / norm / a 32x-decimated field it stores
    the error amplification ||A^-1||_2 ||b|| / ||x|| of an iterative solve (power iteration on A^-H A^-1 with
    the reference's own cached LU), which the default residual tolerance of the GPU solver is derived from.
    Also re-measures the same factor at 300^2 and 1024^2 (key amp_N)."""
    import time
    out = dict(N=2048)
    for N in (300, 1024, 2048):
        for pol in ("TM", "TE"):
            sim = _ref_sim(N, -1e6, 40, 80)
            t0 = time.time()
            F = getattr(sim, "solve_total_field_" + pol)()
            dt = time.time() - t0
            A = getattr(sim, "_A_" + pol)
            b = np.asarray((sim.Q @ A - A @ sim.Q) @ sim.source).ravel()
            lu = getattr(sim, "_LU_" + pol).__self__          # scipy SuperLU object behind factorized()
            rng = np.random.default_rng(1)
            v = rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N)
            v /= np.linalg.norm(v)
            s2 = 0.0
            for _ in range(12):
                w = lu.solve(v)
                w = lu.solve(w, trans="H")
                s2 = np.linalg.norm(w)
                v = w / s2
            amp = np.sqrt(s2) * np.linalg.norm(b) / np.linalg.norm(F)
            out[f"amp_{N}_{pol}"] = amp
            print(N, pol, "solve", round(dt, 1), "s  amp", amp, flush=True)
            if N == 2048:
                out[f"probe_{pol}"] = F[1024, 512]
                out[f"max_{pol}"] = np.abs(F).max()
                out[f"norm_{pol}"] = np.linalg.norm(F)
                out[f"dec_{pol}"] = F[::32, ::32].copy()
                out[f"reference_seconds_{pol}"] = np.float64(dt)
            del sim, lu, A, F
    np.savez_compressed(os.path.join(HERE, "scatter_2048.npz"), **out)


def band_fixture():
    """BASELINE config 3 (example_square_lattice.py: a=1, 40x40, eps_bg=10.2, air hole r=0.4, 200-point
    Gamma-X-M-Gamma path, 5 bands) at path indices 0, 33, 59, 100, 117, 199, and the rectangular
    example (example_rectangular_unitcell.py: 48x64, b=1.4) at 5 indices of its 240-point path."""
    mod = ref_loader.load("Band_Diagram_Solver")
    out = {}
    s = mod.BandDiagramSolver2D(a=1.0, Nx=40, background_er=10.2)
    s.add_circular_inclusion(radius=0.4, er=1.0)
    pts, _ = s.default_rectangular_lattice_path()
    bp, ticks = s.generate_bloch_path(pts, total_points=200)
    idx = np.array([0, 33, 59, 100, 117, 199])
    r = s.compute_band_structure(bp[:, idx], num_bands=5)
    out.update(sq_path=bp, sq_ticks=np.array(ticks), sq_idx=idx, sq_ev_TM=r.eigenvalues["TM"], sq_ev_TE=r.eigenvalues["TE"],
               sq_f_TM=r.frequencies["TM"], sq_f_TE=r.frequencies["TE"])
    s = mod.BandDiagramSolver2D(a=1.0, b=1.4, Nx=48, Ny=64, background_er=11.9)
    s.add_object(lambda X, Y: (np.abs(X) < 0.15) & (Y > -0.2) & (Y < 0.45), er=4.2)
    s.add_circular_inclusion(radius=0.28, center=(-0.15, 0.0), er=1.0)
    pts, _ = s.default_rectangular_lattice_path()
    bp, ticks = s.generate_bloch_path(pts, total_points=240)
    idx = np.array([5, 60, 120, 180, 230])
    r = s.compute_band_structure(bp[:, idx], num_bands=6)
    out.update(rect_path=bp, rect_idx=idx, rect_ev_TM=r.eigenvalues["TM"], rect_ev_TE=r.eigenvalues["TE"],
               rect_f_TM=r.frequencies["TM"], rect_f_TE=r.frequencies["TE"])
    np.savez_compressed(os.path.join(HERE, "band_golden.npz"), **out)


def mode_cases():
    """Problem definitions shared by the fixture generator and the tests (name -> builder(cls))."""
    def ridge(cls):          # BASELINE config 2: example_ridge_dielectric_waveguide.py
        s = cls(50e9, 24e-3, 16e-3, 240, 160, 5)
        s.add_rectangle((3.0, 4.0, 5.0), 1, (0, 240), (60, 80))
        s.add_rectangle(6.0, 1, (100, 140), (80, 100))
        s.add_UPML(pml_width=30, n=3, sigma_max=1, direction="x")
        return s

    def strip(cls):          # small lossy microstrip-like case: PEC strip + ground, lossy substrate
        s = cls(30e9, 12e-3, 8e-3, 60, 40, 3)
        s.add_rectangle(4 + 1j, 1, (2e-3, 10e-3), (2e-3, 3e-3))
        s.add_pec((5e-3, 7e-3), (3e-3, 3.2e-3))
        s.add_pec((0, 60), (9, 10))
        return s

    def shapes(cls):         # circle + triangle + impedance sheet + PMC wall + y PML
        s = cls(40e9, 10e-3, 10e-3, 50, 50, 4)
        s.add_circle(5.0, 1.0, (5e-3, 5e-3), 2e-3, 0.8e-3)
        s.add_triangle((2.0, 2.5, 3.0), 1.2, (1e-3, 1e-3), (4e-3, 1.5e-3), (2e-3, 3.5e-3))
        s.add_impedance_surface(50 + 20j, 40, orientation="x")
        s.add_pmc((0, 2), (0, 50), components=("yy", "zz"))
        s.add_pml(pml_width=8, sigma_max=2, direction="y")
        return s
    return dict(ridge=ridge, strip=strip, shapes=shapes)


def mode_fixture():
    mod = ref_loader.load("Mode_Solver_2D")
    out = {}
    for name, build in mode_cases().items():
        s = build(mod.ModeSolver2D)
        s.solve()
        out[f"{name}_eigenvalues"] = s.eigenvalues
        out[f"{name}_neff"] = s.neff
        out[f"{name}_sigma"] = np.array(s._resolve_eigs_guess(None))
        for fld in ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz"):
            a = getattr(s, fld)
            out[f"{name}_{fld}_norm"] = np.array([np.linalg.norm(a[:, :, m]) for m in range(s.num_modes)])
            out[f"{name}_{fld}_dec"] = a[::7, ::5, :].copy()
    np.savez_compressed(os.path.join(HERE, "mode_golden.npz"), **out)


def periodic3d_cases():
    def tiny(cls):
        s = cls(Nx=8, Ny=7, Nz=6, x_range=4e-3, y_range=3.5e-3, z_range=3e-3, freq=25e9, num_modes=2,
                sigma_guess=0, tol=1e-10, ncv=14)
        s.add_pec((1e-3, 3e-3), (1e-3, 1.5e-3), (0, 1.5e-3))
        s.add_block((6, 5, 4), 1, (1e-3, 3e-3), (0, 1e-3), (0, 6), subpixels=4)
        s.add_UPML(['+y'], width=2, max_loss=3, n=3)
        return s, dict(method="refined", max_restarts=6)

    def sweep21(cls):        # first point of Periodic_3D_Dispersion.py (20x18x16, 21 GHz)
        f = 21e9
        k0 = 2 * np.pi * f / 3e8
        s = cls(Nx=20, Ny=18, Nz=16, x_range=6e-3, y_range=6e-3, z_range=8e-3, freq=f, num_modes=2,
                sigma_guess=1j * k0 * (0.12 * 21 - 2.7), tol=1e-2, ncv=30)
        s.add_pec(slice(0, 3), slice(10, 11), slice(0, 16))
        s.add_pec(slice(5, 8), slice(10, 11), slice(0, 16))
        s.add_block(6, 1, slice(3, 17), slice(11, 17), slice(0, 16), subpixels=8)
        s.add_pec(slice(0, 20), slice(17, 18), slice(0, 16))
        s.add_UPML(['+y'], width=8, max_loss=10, n=3)
        return s, dict(method="refined", max_restarts=8)

    def small_eigs(cls):
        # NB: the reference's method="eigs" (ARPACK) is not reproducible on this pencil (singular B, highly
        # non-normal: it returns start-vector dependent pseudo-eigenvalues near sigma),
        # so the golden values come from its "refined" method; the GPU "eigs" path (Krylov-Schur) is
        # tested against them.
        f = 21e9
        k0 = 2 * np.pi * f / 3e8
        s = cls(Nx=10, Ny=9, Nz=8, x_range=6e-3, y_range=6e-3, z_range=8e-3, freq=f, num_modes=2,
                sigma_guess=1j * k0 * (0.12 * 21 - 2.7), tol=1e-10, ncv=30)
        s.add_block(6, 1, slice(2, 8), slice(5, 8), slice(0, 8), subpixels=8)
        s.add_pmc(slice(0, 10), slice(0, 1), slice(0, 8))
        s.add_pec(slice(0, 10), slice(8, 9), slice(0, 8))
        s.add_UPML(['+y', '-x'], width=3, max_loss=10, n=3)
        return s, dict(method="refined", max_restarts=8)
    return dict(tiny=tiny, sweep21=sweep21, small_eigs=small_eigs)


def periodic3d_cfg4(cls):
    """BASELINE config 4: example_image_guide_leaky_wave_antenna.py:7-28 verbatim (30x30x32, n = 119 040)."""
    s = cls(Nx=30, Ny=30, Nz=32, x_range=6e-3, y_range=6e-3, z_range=8e-3, freq=22e9, num_modes=2, sigma_guess=0,
            tol=0.1, ncv=30)
    s.add_pec((1.5e-3, 4.5e-3), (1.5e-3, 1.55e-3), (0, 1.5e-3))
    s.add_block(6, 1, (1.5e-3, 4.5e-3), (0, 1.5e-3), (0, 32), subpixels=8)
    s.add_UPML(['+y'], width=10, max_loss=5, n=3)
    return s, dict(method="refined", max_restarts=8)


def periodic3d_cfg4_fixture():
    """~40 min of SuperLU on the host: kept in its own file so the other goldens regenerate fast."""
    import time
    mod = ref_loader.load("Periodic_Solver_3D")
    s, kw = periodic3d_cfg4(mod.PeriodicModeSolver3D)
    t0 = time.time()
    s.solve(**kw)
    out = dict(eigenvalues=s.eigenvalues, gammas=s.gammas, residuals=s.refined_residuals,
               restarts=np.int64(s.refined_restarts), reference_seconds=np.float64(time.time() - t0))
    for fld in ("Ex", "Hy"):
        out[f"{fld}_norm"] = np.array([np.linalg.norm(s.fields[fld][m]) for m in range(s.num_modes)])
        out[f"{fld}_dec"] = s.fields[fld][:, ::3, ::3, ::4].copy()
    np.savez_compressed(os.path.join(HERE, "periodic3d_cfg4_golden.npz"), **out)


def periodic3d_fixture():
    mod = ref_loader.load("Periodic_Solver_3D")
    out = {}
    for name, build in periodic3d_cases().items():
        s, kw = build(mod.PeriodicModeSolver3D)
        s.solve(**kw)
        out[f"{name}_eigenvalues"] = s.eigenvalues
        out[f"{name}_gammas"] = s.gammas
        if s.refined_residuals is not None:
            out[f"{name}_residuals"] = s.refined_residuals
            out[f"{name}_restarts"] = np.int64(s.refined_restarts)
        for fld in ("Ex", "Hy"):
            out[f"{name}_{fld}_norm"] = np.array([np.linalg.norm(s.fields[fld][m]) for m in range(s.num_modes)])
            out[f"{name}_{fld}_dec"] = s.fields[fld][:, ::3, ::3, ::2].copy()
    np.savez_compressed(os.path.join(HERE, "periodic3d_golden.npz"), **out)


def periodic2d_cases():
    """Problem definitions of the 2-D periodic solver shared by the fixture generator and the tests."""
    def antenna(cls):        # Periodic_Solver_2D/example_surface_wave_antenna.py:3-22 verbatim (TM, 200 x 80, n = 32 000)
        s = cls("TM", freq=20e9, x_range=10e-3, z_range=8e-3, Nx=200, Nz=80, num_modes=8, guess=0, mode_filter=False)
        s.add_pec((2.27e-3, 2.28e-3), (1.0e-3, 2.0e-3))
        s.add_pec((0.9e-3, 1.0e-3), (0, 80))
        s.add_rectangle(10.2, 1, (1.0e-3, 2.27e-3), (0, 80), subpixels=16)
        s.add_pml(pml_width=50, n=3, sigma_max=5, direction="x+")
        return s

    def slab_te(cls):        # TE modes of a dielectric slab on a PEC-like ground with a PMC-like strip; both PMLs
        s = cls("TE", freq=30e9, x_range=6e-3, z_range=4e-3, Nx=60, Nz=16, num_modes=4, guess=0)
        s.add_pec((0, 3), (0, 16))
        s.add_rectangle((9.8, 9.8, 6.0), 1, (0.3e-3, 1.7e-3), (0, 16), subpixels=8)
        s.add_pmc((30, 32), (4, 9), components=("xx", "zz"))
        s.add_pml(pml_width=15, n=3, sigma_max=4, direction="x+")
        s.guess = (-0.08 + 0.87j) * s.k0          # next to the one pair ARPACK converges here (neff = -0.0811 + 0.8712j)
        return s

    def grating_tm(cls):     # TM, periodic dielectric teeth (z-dependent cell), small: dense shifted solve
        s = cls("TM", freq=25e9, x_range=5e-3, z_range=3e-3, Nx=40, Nz=12, num_modes=4, guess=0)
        s.add_pec((0, 2), (0, 12))
        s.add_rectangle(6.15, 1, (0.2e-3, 1.2e-3), (0, 12), subpixels=8)
        s.add_rectangle(6.15, 1, (1.2e-3, 1.7e-3), (0.4e-3, 1.6e-3), subpixels=8)
        s.add_pml(pml_width=12, n=3, sigma_max=5, direction="x+")
        s.guess = (-0.02 + 1.39j) * s.k0          # next to neff = -0.0183 + 1.3869j
        return s
    return dict(antenna=antenna, slab_te=slab_te, grating_tm=grating_tm)


def periodic2d_fixture():
    """Eigenvalues of the unmodified reference with the relative residual of every returned pair in its own pencil:
    the parity tests compare the pairs ARPACK actually converged (see scripts/periodic2d_reproducibility.py)."""
    from oracle import periodic2d_oracle as po
    mod = ref_loader.load("Periodic_Solver_2D")
    out = {}
    for name, build in periodic2d_cases().items():
        s = build(mod.PeriodicModeSolver2D)
        s.solve()
        m = s._effective_materials_and_masks()
        A, B = (s._build_tm_system(m[0], m[2], m[4]) if s.polarization == "TM" else s._build_te_system(m[1], m[3], m[5]))
        free = s._free_mask(m[6], m[7], m[9], m[10])
        A, B = A[free, :][:, free], B[free, :][:, free]
        out[f"{name}_eigenvalues"] = s.eigenvalues
        out[f"{name}_neff"] = s.neff
        out[f"{name}_residuals"] = po.relative_residuals(A, B, s.eigenvalues, s.eigenvectors[free, :])
        conv = np.flatnonzero(out[f"{name}_residuals"] <= 1e-9)
        out[f"{name}_converged"] = conv
        first, second = ("Ex", "Hy") if s.polarization == "TM" else ("Hx", "Ey")
        for fld in (first, second):                      # decimated fields of the converged pairs only
            arr = getattr(s, fld)
            shape = getattr(s, "shape_" + fld.lower())
            out[f"{name}_{fld}_dec"] = np.array([arr[:, k].reshape(shape, order="F")[::3, ::2] for k in conv])
    np.savez_compressed(os.path.join(HERE, "periodic2d_golden.npz"), **out)
    for name in periodic2d_cases():
        print(name, out[f"{name}_eigenvalues"], out[f"{name}_residuals"])


if __name__ == "__main__":
    if "--p2d-only" in sys.argv:
        periodic2d_fixture()
        sys.exit(0)
    if "--scatter-1024" in sys.argv:
        scatter_1024()
        sys.exit(0)
    if "--scatter-2048" in sys.argv:
        scatter_2048()
        sys.exit(0)
    if "--p3d-cfg4" in sys.argv:
        periodic3d_cfg4_fixture()
        sys.exit(0)
    if "--p3d-only" in sys.argv:
        periodic3d_fixture()
        sys.exit(0)
    if "--mode-only" in sys.argv:
        mode_fixture()
        sys.exit(0)
    if "--band-only" in sys.argv:
        band_fixture()
        sys.exit(0)
    yee_fixture()
    band_fixture()
    mode_fixture()
    periodic3d_fixture()
    periodic2d_fixture()
    scatter_small()
    scatter_cfg1()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
