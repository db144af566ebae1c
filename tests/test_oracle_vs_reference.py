"""Pins the oracle (oracle/*.py) against the UNMODIFIED reference imported from /root/reference.
Runs in the build container only (skipped where the reference tree is absent)."""
import itertools

import numpy as np
import pytest

from conftest import same_sparse
from oracle import ref_loader, scattering_oracle as so, yee_oracle

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("folder", ["Scattering", "Band_Diagram_Solver"])
def test_yeeder2d_bitwise(folder):
    ref = ref_loader.ref_yeeder2d(folder)
    rng = np.random.default_rng(1)
    n = 0
    for Nx, Ny in itertools.product([1, 2, 3, 4, 7, 16, 33], repeat=2):
        for bc in itertools.product([0, 1], repeat=2):
            for kinc in (None, [0.3, 0.7], rng.standard_normal(2)):
                res = [0.5, 0.25] if kinc is None else list(rng.uniform(0.01, 1, 2))
                R = ref([Nx, Ny], res, list(bc), kinc)
                O = yee_oracle.yeeder2d([Nx, Ny], res, list(bc), kinc)
                for r, o in zip(R, O):
                    assert same_sparse(r, o), (Nx, Ny, bc, kinc)
                    n += 1
    assert n == 7 * 7 * 4 * 3 * 4


@pytest.mark.parametrize("N,eps,pw,mask", [(96, 4.0, 16, 24), (128, -1e6, 20, 40)])
def test_scattering_pieces(N, eps, pw, mask):
    mod = ref_loader.load("Scattering_Solver_2D")
    f0 = 10e9
    lam = 299792458 / f0
    sim = mod.FDFD2DScatteringSolver(frequency=f0, x_range=N / 50 * lam, y_range=N / 50 * lam, Nx=N, Ny=N)
    sim.add_object(er_tensor=eps, mr_tensor=1.0, region_mask=(sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=pw, sigma_max=10)
    sim.add_mask(value=mask)
    sim.add_source(src_type="plane_wave", angle_deg=45.0, polarization="TE")
    p = so.cylinder_problem(N, eps_r=eps, pml_width=pw, mask=mask)
    for k in p.mats:
        assert np.array_equal(p.mats[k], getattr(sim, k))
    assert np.array_equal(p.source, sim.source)
    assert np.array_equal(p.q, sim.Q.diagonal())
    rng = np.random.default_rng(0)
    x = rng.standard_normal(p.N) + 1j * rng.standard_normal(p.N)
    for pol in ("TE", "TM"):
        F = getattr(sim, "solve_total_field_" + pol)()
        Aref = getattr(sim, "_A_" + pol)
        A = so.build_A(p, pol)
        assert A.format == Aref.format and (A != Aref).nnz == 0
        bref = np.asarray((sim.Q @ Aref - Aref @ sim.Q) @ sim.source).ravel()
        assert np.array_equal(so.rhs(A, p), bref)
        b2 = so.rhs_matrix_free(p, pol)
        # cancellation: both terms of Q(As) - A(Qs) are O(|A s|) (1e6 inside the eps=-1e6 cylinder)
        assert np.linalg.norm(b2 - bref) <= 1e-14 * np.linalg.norm(Aref @ sim.source)
        y = so.apply_stencil(pol, *so.stencil_coefficients(p, pol), x).ravel()
        yr = Aref @ x
        assert np.linalg.norm(y - yr) <= 1e-14 * np.linalg.norm(yr)
        Fo = so.solve_total_field(p, pol, A)
        assert np.linalg.norm(Fo - F) <= 1e-12 * np.linalg.norm(F)


def test_point_source_matches():
    mod = ref_loader.load("Scattering_Solver_2D")
    sim = mod.FDFD2DScatteringSolver(10e9, 0.06, 0.05, 40, 30)
    p = so.make_problem(10e9, 0.06, 0.05, 40, 30)
    for pol in ("TE", "TM"):
        sim.add_source("point", polarization=pol, location=(0.001, -0.002), amplitude=2.0)
        so.add_source(p, "point", polarization=pol, location=(0.001, -0.002), amplitude=2.0)
        assert np.array_equal(sim.source, p.source)


def test_band_oracle_matches_reference():
    """Band-diagram oracle vs the unmodified reference class on a small cell (path, materials, eigenvalues)."""
    from oracle import band_oracle as bo
    mod = ref_loader.load("Band_Diagram_Solver")
    s = mod.BandDiagramSolver2D(a=1.0, Nx=20, Ny=16, b=0.9, background_er=10.2)
    s.add_circular_inclusion(radius=0.3, er=1.0)
    p = bo.BandProblem(1.0, 20, 16, b=0.9, background_er=10.2)
    p.add_circle(0.3, er=1.0)
    assert np.array_equal(s.ER2, p.ER2)
    pts, _ = s.default_rectangular_lattice_path()
    bp, ticks = s.generate_bloch_path(pts, total_points=23)
    bp2, ticks2 = bo.bloch_path(bo.gamma_x_m_gamma(1.0, 0.9), 23)
    assert np.array_equal(bp, bp2) and list(ticks) == list(ticks2)
    sub = bp[:, [1, 7, 15]]
    r = s.compute_band_structure(sub, num_bands=4)
    o = bo.band_structure(p, sub, 4)
    for pol in ("TE", "TM"):
        assert np.max(np.abs(r.eigenvalues[pol] - o[pol][0])) <= 1e-9 * np.abs(o[pol][0]).max()
        assert np.max(np.abs(r.frequencies[pol] - o[pol][1])) <= 1e-10


def _mode_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.mode_cases()


@pytest.mark.parametrize("case", ["strip", "shapes", "ridge"])
def test_mode_host_logic_and_oracle(case):
    """(1) the drop-in class's host-side pre-processing (sub-pixel geometry, UPML, PEC/PMC masks,
    averaging, staggered operators) is bit-identical to the reference's; (2) the oracle's hot-path
    restatement reproduces the reference eigenvalues."""
    import fdfd_b200
    from oracle import mode_oracle as mo
    build = _mode_cases()[case]
    ref = build(ref_loader.load("Mode_Solver_2D").ModeSolver2D)
    mine = build(fdfd_b200.ModeSolver2D)
    names = [f"cell_{p}_r_{c}" for p in ("eps", "mu") for c in ("xx", "yy", "zz")]
    names += [f"{p}_r_{c}" for p in ("eps", "mu") for c in ("xx", "yy", "zz")]
    names += [f"{p}_{c}_mask" for p in ("pec", "pmc") for c in ("xx", "yy", "zz")] + ["material_no_average_mask"]
    for n in names:
        assert np.array_equal(getattr(ref, n), getattr(mine, n), equal_nan=True), n
    assert ref._resolve_eigs_guess(None) == mine._resolve_eigs_guess(None)
    # the eight staggered operators: the product assembles them on the GPU (K1b; tests/test_stagger_gpu.py
    # byte-compares it with this oracle), here the oracle is pinned against the reference
    D = mo.operators(ref.Nx, ref.Ny, ref.dx_normalized, ref.dy_normalized)
    for a, name in zip(ref._yeeder2d(), ("d_ez_ex", "d_ez_ey", "d_ey_hz", "d_ex_hz", "d_hy_ez", "d_hx_ez", "d_hz_hx", "d_hz_hy")):
        b = D[name]
        assert a.shape == b.shape and a.format == b.format, name
        assert a.indptr.tobytes() == b.indptr.tobytes() and a.indices.tobytes() == b.indices.tobytes(), name
        assert a.data.tobytes() == b.data.tobytes(), name
    rm, mm = ref._effective_materials_and_masks(), mine._effective_materials_and_masks()
    for k in rm[0]:
        assert np.array_equal(rm[0][k], mm[0][k]), k
    for a, b in zip(rm[1:], mm[1:]):
        assert np.array_equal(a, b)
    if case == "ridge":
        return                      # the 77k-unknown eigs call is covered by the golden fixture
    ref.solve()
    cell = {f"{p}_{c}": getattr(ref, f"cell_{p}_r_{c}") for p in ("eps", "mu") for c in ("xx", "yy", "zz")}
    pec = {c: getattr(ref, f"pec_{c}_mask") for c in ("xx", "yy", "zz")}
    pmc = {c: getattr(ref, f"pmc_{c}_mask") for c in ("xx", "yy", "zz")}
    o = mo.solve(cell, ref.material_no_average_mask, pec, pmc, ref.Nx, ref.Ny, ref.dx_normalized, ref.dy_normalized,
                 ref.num_modes, ref._resolve_eigs_guess(None), ref._has_lossy_material())
    assert np.max(np.abs(o["eigenvalues"] - ref.eigenvalues)) <= 1e-10 * np.abs(ref.eigenvalues).max()
    assert np.max(np.abs(o["neff"] - ref.neff)) <= 1e-10
    # (the factored operator the GPU path uses, P_e Q_e = Omega, is checked in tests/test_stagger_gpu.py: its
    # staggered operators come from the GPU assembly kernel)


def _p3d_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.periodic3d_cases()


@pytest.mark.parametrize("case", ["tiny", "small_eigs", "sweep21"])
def test_periodic3d_host_logic_and_oracle(case):
    """Drop-in 3-D class: cell tensors, Yee-location materials, the sixteen operators and the pencil
    (A, B) are bit-identical to the reference's; the oracle's refined Arnoldi reproduces the
    reference eigenvalues (tiny case)."""
    import fdfd_b200
    from oracle import periodic3d_oracle as po
    build = _p3d_cases()[case]
    ref, kw = build(ref_loader.load("Periodic_Solver_3D").PeriodicModeSolver3D)
    mine, _ = build(fdfd_b200.PeriodicModeSolver3D)
    for n in ("Erxx", "Eryy", "Erzz", "Mrxx", "Mryy", "Mrzz"):
        assert np.array_equal(getattr(ref, f"cell_{n}_3D"), getattr(mine, f"cell_{n}_3D")), n
        assert np.array_equal(getattr(ref, f"{n}_3D"), getattr(mine, f"{n}_3D")), n
    assert np.array_equal(ref.material_no_average_mask, mine.material_no_average_mask)
    ops = ["DEX_EZ_TO_EX", "DEY_EZ_TO_EY", "DEX_EY_TO_HZ", "DEY_EX_TO_HZ", "DHX_HY_TO_EZ", "DHY_HX_TO_EZ",
           "DHX_HZ_TO_HX", "DHY_HZ_TO_HY", "DEZ_EX", "DEZ_EY", "DHZ_HX", "DHZ_HY", "AZ_EX", "AZ_EY", "AZ_HX", "AZ_HY"]
    D = po.operators(ref.Nx, ref.Ny, ref.Nz, ref.dx, ref.dy, ref.dz)
    for name in ops:
        r = getattr(ref, name)
        other = D[name]         # the product's GPU-assembled operators: tests/test_stagger_gpu.py
        assert r.shape == other.shape and r.format == other.format, name
        assert r.indptr.tobytes() == other.indptr.tobytes() and r.indices.tobytes() == other.indices.tobytes(), name
        assert r.data.tobytes() == other.data.tobytes(), name
    Ar, Br = ref._build_eigen_matrices()
    cell_o = {k: getattr(ref, f"cell_{k[0].upper()}{k[1:]}_3D") for k in ("erxx", "eryy", "erzz", "mrxx", "mryy", "mrzz")}
    Ao, Bo = po.pencil(po.materials_on_fields(cell_o, ref.material_no_average_mask), D, ref.omega)
    assert abs(Ar - Ao).max() <= 1e-12 * abs(Ar).max() and abs(Br - Bo).max() == 0
    if case != "tiny":
        return
    ref.solve(**kw)
    cell = {k: getattr(ref, f"cell_{k[0].upper()}{k[1:]}_3D") for k in ("erxx", "eryy", "erzz", "mrxx", "mryy", "mrzz")}
    mats = po.materials_on_fields(cell, ref.material_no_average_mask)
    A, B = po.pencil(mats, D, ref.omega)
    assert (A != Ar).nnz == 0 and (B != Br).nnz == 0
    vals, vecs, res, rs = po.refined_shift_invert_arnoldi(A, B, ref.sigma_guess, ref.num_modes, ref.tol, ncv=ref.ncv,
                                                          max_restarts=kw["max_restarts"], seed=0)
    assert np.max(np.abs(vals - ref.eigenvalues)) <= 1e-9 * np.abs(ref.eigenvalues).max()
    assert rs == ref.refined_restarts and np.allclose(res, ref.refined_residuals, rtol=1e-6)


def _p2d_cases():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.periodic2d_cases()


@pytest.mark.parametrize("case", ["antenna", "slab_te", "grating_tm"])
def test_periodic2d_host_logic_and_oracle(case):
    """Drop-in 2-D periodic class: cell tensors, field-location materials, masks, the twelve operators and the
    pencil (A, B, free mask) are bit-identical to the reference's; error behaviour of the region helpers too."""
    import fdfd_b200
    from oracle import periodic2d_oracle as po
    build = _p2d_cases()[case]
    ref = build(ref_loader.load("Periodic_Solver_2D").PeriodicModeSolver2D)
    mine = build(fdfd_b200.PeriodicModeSolver2D)
    assert mine.k0 == ref.k0 and mine.guess == ref.guess and mine.shape_ez == ref.shape_ez
    comps = [f"{p}_r_{c}" for p in ("eps", "mu") for c in ("xx", "yy", "zz")]
    for n in comps:
        assert np.array_equal(getattr(ref, "cell_" + n), getattr(mine, "cell_" + n)), n
        assert np.array_equal(getattr(ref, n), getattr(mine, n)), n
    assert np.array_equal(ref.material_no_average_mask, mine.material_no_average_mask)
    assert ref._pec_regions == mine._pec_regions and ref._pmc_regions == mine._pmc_regions
    D = po.operators(ref.Nx, ref.Nz, ref.dx, ref.dz)
    for name in fdfd_b200.PeriodicModeSolver2D._OPERATORS:
        r, other = getattr(ref, name), D[name]          # the product's GPU-assembled operators: tests/test_periodic2d_gpu.py
        assert r.shape == other.shape and r.format == other.format, name
        assert r.indptr.tobytes() == other.indptr.tobytes() and r.indices.tobytes() == other.indices.tobytes(), name
        assert r.data.tobytes() == other.data.tobytes(), name
        setattr(mine, name, other)                      # stand in for the GPU assembly on this CPU box
    mr, mm = ref._effective_materials_and_masks(), mine._effective_materials_and_masks()
    assert len(mr) == len(mm) == 12
    for a, b in zip(mr, mm):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    if ref.polarization == "TM":
        Ar, Br = ref._build_tm_system(mr[0], mr[2], mr[4])
    else:
        Ar, Br = ref._build_te_system(mr[1], mr[3], mr[5])
    free = ref._free_mask(mr[6], mr[7], mr[9], mr[10])
    Am, Bm, fm = mine._pencil()
    assert np.array_equal(free, fm)
    Ar, Br = Ar[free, :][:, free], Br[free, :][:, free]
    assert (Ar != Am).nnz == 0 and (Br != Bm).nnz == 0
    cell = {f"{p}_{c}": getattr(ref, f"cell_{p}_r_{c}") for p in ("eps", "mu") for c in ("xx", "yy", "zz")}
    Ao, Bo = po.pencil(ref.polarization, po.materials_on_fields(cell, ref.material_no_average_mask), D, ref.omega)
    assert (Ao[free, :][:, free] != Ar).nnz == 0 and (Bo[free, :][:, free] != Br).nnz == 0
    assert np.array_equal(mine._unknown_planes().shape, free.shape)
    for bad in (lambda s: s.add_rectangle(2, 1, (0.0, 1.0), (0, 4)), lambda s: s.add_pec((5, 5), (0, 4)),
                lambda s: s.add_pmc((0, 2), (0, 3), components=("xy",)), lambda s: s.add_pml(direction="z"),
                lambda s: s.add_rectangle(2, 1, (0, 4), (0, 4), subpixels=0), lambda s: s.add_pec("ab", (0, 4))):
        msgs = []
        for obj in (ref, mine):
            with pytest.raises(ValueError) as e:
                bad(obj)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1]


@pytest.mark.parametrize("N,pw,direction", [(120, 20, "both"), (97, 11, "x"), (64, 9, "y"), (40, 20, "both")])
def test_scattering_dropin_host_arrays_bitwise(N, pw, direction):
    """Host side of the drop-in class (frame-only UPML scaling, broadcast grid, lazy identity mask) against
    the unmodified reference class: material tensors and source bit-identical, including overlapping layers."""
    import fdfd_b200
    R = ref_loader.load("Scattering_Solver_2D").FDFD2DScatteringSolver
    f0 = 10e9
    lam = 299792458 / f0
    args = (f0, N / 50 * lam, (N + 3) / 50 * lam, N, N + 3)
    sim, ref = fdfd_b200.FDFD2DScatteringSolver(*args), R(*args)
    for s in (sim, ref):
        s.add_object([4.0 - 0.3j, 2.0, 3.0 + 1j], 1.5, (s.X ** 2 + s.Y ** 2) <= (0.5 * lam) ** 2)
        s.add_UPML(pml_width=pw, sigma_max=10, direction=direction)
        s.add_source("plane_wave", angle_deg=33.0)
    for name in ("ERxx", "ERyy", "ERzz", "MRxx", "MRyy", "MRzz", "source"):
        assert np.array_equal(getattr(sim, name), getattr(ref, name)), name
    assert (sim.Q != ref.Q).nnz == 0
    assert np.array_equal(sim.X, ref.X) and np.array_equal(sim.Y, ref.Y)
    sim.add_mask(7)
    ref.add_mask(7)
    assert (sim.Q != ref.Q).nnz == 0
