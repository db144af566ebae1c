"""GPU parity of the Krylov stack (K3-K5) and of the drop-in scattering solver against the
oracle's direct (SuperLU) solve.  Bars from BASELINE.json: residual <= 1e-8, field relative
L2 error <= 1e-6 in complex128."""
import numpy as np
import pytest

from oracle import scattering_oracle as so

pytestmark = pytest.mark.gpu


def _lossy_problem(N, pol):
    """A well-conditioned (strongly lossy) system for the plain Krylov checks."""
    rng = np.random.default_rng(3)
    ca = 1.0 + 0.1 * rng.standard_normal((N, N)) + 0.05j
    cb = 1.0 + 0.1 * rng.standard_normal((N, N)) - 0.05j
    cd = (-300.0 + 40.0j) * np.ones((N, N))
    return ca, cb, cd, 63.0, 63.0


@pytest.mark.parametrize("method,precond", [("bicgstab", "none"), ("bicgstab", "jacobi"),
                                            ("fgmres", "none"), ("fgmres", "jacobi")])
@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_krylov_vs_direct(method, precond, pol):
    import torch
    import scipy.sparse.linalg as spla
    import fdfd_b200
    from test_helm_gpu import _explicit
    N = 48
    ca, cb, cd, sx, sy = _lossy_problem(N, pol)
    M = _explicit(pol, N, N, ca, cb, cd, sx, sy, (0, 0), None)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N)
    xref = spla.spsolve(M.tocsc(), b)
    A = fdfd_b200.HelmholtzOperator2D(pol, N, N, ca, cb, cd, sx, sy)
    x, info = A.solve(torch.from_numpy(b).cuda(), method=method, precond=precond, rtol=1e-10, maxit=2000)
    x = x.cpu().numpy()
    assert info["relres"] <= 1e-10
    assert np.linalg.norm(M @ x - b) <= 1.5e-10 * np.linalg.norm(b)
    assert np.linalg.norm(x - xref) <= 1e-8 * np.linalg.norm(xref)
    # host-buffer entry point gives the same answer
    xh, _ = A.solve(b, method=method, precond=precond, rtol=1e-10, maxit=2000)
    assert np.linalg.norm(xh - xref) <= 1e-8 * np.linalg.norm(xref)


def test_noconvergence_is_loud():
    import torch
    import fdfd_b200
    N = 48
    ca, cb, cd, sx, sy = _lossy_problem(N, "TE")
    A = fdfd_b200.HelmholtzOperator2D("TE", N, N, ca, cb, cd, sx, sy)
    b = torch.ones(N * N, dtype=torch.complex128, device="cuda")
    with pytest.raises(fdfd_b200.NoConvergence):
        A.solve(b, method="bicgstab", precond="none", rtol=1e-14, maxit=2)


@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_scattering_small_vs_golden(golden, pol):
    """120x120 reference problem (eps_r=4 cylinder, UPML 20, mask 30): drop-in class vs the field
    the unmodified reference produced."""
    import fdfd_b200
    g = golden("scatter_small.npz")
    N = int(g["N"])
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(float(g["eps"]), 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=int(g["pml"]), sigma_max=10)
    sim.add_mask(int(g["mask"]))
    sim.add_source("plane_wave", angle_deg=45.0, polarization="TE")
    F = getattr(sim, "solve_total_field_" + pol)()
    ref = g[f"F_{pol}"]
    assert F.shape == (N, N)
    assert sim.solver_info[pol]["relres"] <= 1e-8
    assert np.linalg.norm(F - ref) <= 1e-6 * np.linalg.norm(ref)
    # right-hand side built with two stencil applies == reference's (QA - AQ) src
    b = sim._rhs(sim.operator(pol)).cpu().numpy()
    assert np.linalg.norm(b - g[f"b_{pol}"]) <= 1e-12 * np.linalg.norm(g[f"b_{pol}"])


@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_scattering_cfg1(golden, pol):
    """BASELINE config 1 (example_scattering_by_cylinder.py: 300x300, eps_r=-1e6, UPML 40, mask 80)."""
    import fdfd_b200
    g = golden("scatter_cfg1.npz")
    N = 300
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, 6 * lam, 6 * lam, N, N)
    sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=40, sigma_max=10)
    sim.add_mask(80)
    sim.add_source("plane_wave", angle_deg=45.0, polarization="TE")
    F = getattr(sim, "solve_total_field_" + pol)()
    assert sim.solver_info[pol]["relres"] <= 1e-8
    dec = g[f"dec_{pol}"]
    assert np.linalg.norm(F[::10, ::10] - dec) <= 1e-6 * np.linalg.norm(dec)
    assert abs(F[150, 75] - g[f"probe_{pol}"]) <= 1e-6 * abs(g[f"max_{pol}"])
    assert abs(np.linalg.norm(F) - g[f"norm_{pol}"]) <= 1e-6 * g[f"norm_{pol}"]
    # same problem through the oracle's direct solve, full field
    p = so.cylinder_problem(N)
    ref = so.solve_total_field(p, pol)
    assert np.linalg.norm(F - ref) <= 1e-6 * np.linalg.norm(ref)


def test_repeated_solves_reuse_preconditioner(golden):
    """Second solve on the same operator (new source) reuses the cached multigrid hierarchy, like the
    reference reuses its LU (`reuse_factorisation=True`); results must still match the oracle."""
    import time
    import fdfd_b200
    N = 192
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(4.0, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=20, sigma_max=10)
    sim.add_mask(40)
    p = so.cylinder_problem(N, eps_r=4.0, pml_width=20, mask=40)
    times = []
    for angle, reuse in ((45.0, True), (10.0, True), (70.0, False)):
        sim.add_source("plane_wave", angle_deg=angle)
        so.add_source(p, "plane_wave", angle_deg=angle)
        t = time.perf_counter()
        F = sim.solve_total_field_TE(reuse_factorisation=reuse)
        times.append(time.perf_counter() - t)
        ref = so.solve_total_field(p, "TE")
        assert np.linalg.norm(F - ref) <= 1e-6 * np.linalg.norm(ref), angle
    assert sim.solver_info["TE"]["relres"] <= 1e-8


@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_complex64_solve(pol):
    """complex64 path (40 B/pt operator, fp32 multigrid + FGMRES): residual 1e-4, field within 1e-2
    (reduced-precision option; the parity bars of BASELINE.json are stated for complex128)."""
    import fdfd_b200
    N = 160
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N, dtype="complex64")
    sim.add_object(4.0, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=16, sigma_max=10)
    sim.add_mask(32)
    sim.add_source("plane_wave", angle_deg=45.0)
    sim.solver_options.update(rtol=1e-4)
    F = getattr(sim, "solve_total_field_" + pol)()
    p = so.cylinder_problem(N, eps_r=4.0, pml_width=16, mask=32)
    ref = so.solve_total_field(p, pol)
    assert sim.solver_info[pol]["relres"] <= 1e-4
    assert np.linalg.norm(F - ref) <= 1e-2 * np.linalg.norm(ref)


@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_mixed_precision_preconditioner(pol):
    """complex64 multigrid cycle inside the complex128 FGMRES: same 1e-8 residual / 1e-6 field bars."""
    import fdfd_b200
    N = 192
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=20, sigma_max=10)
    sim.add_mask(40)
    sim.add_source("plane_wave", angle_deg=45.0)
    sim.solver_options.update(mg_single=1)
    F = getattr(sim, "solve_total_field_" + pol)()
    its_mixed = sim.solver_info[pol]["iterations"]
    p = so.cylinder_problem(N, eps_r=-1e6, pml_width=20, mask=40)
    ref = so.solve_total_field(p, pol)
    assert sim.solver_info[pol]["relres"] <= 1e-8
    assert np.linalg.norm(F - ref) <= 1e-6 * np.linalg.norm(ref)
    sim2 = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim2.add_object(-1e6, 1.0, (sim2.X ** 2 + sim2.Y ** 2) <= (0.5 * lam) ** 2)
    sim2.add_UPML(pml_width=20, sigma_max=10)
    sim2.add_mask(40)
    sim2.add_source("plane_wave", angle_deg=45.0)
    getattr(sim2, "solve_total_field_" + pol)()
    assert its_mixed <= 1.3 * sim2.solver_info[pol]["iterations"] + 3


@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_compressed_krylov_basis(pol):
    """FGMRES with the basis V stored in complex64 (arithmetic, corrections and residual in complex128):
    same 1e-8 residual / 1e-6 field bars, iteration count within a few of the complex128 basis."""
    import fdfd_b200
    N = 192
    f0 = 10e9
    lam = 299792458 / f0
    its = {}
    for bs in (0, 1):
        sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
        sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
        sim.add_UPML(pml_width=20, sigma_max=10)
        sim.add_mask(40)
        sim.add_source("plane_wave", angle_deg=45.0)
        sim.solver_options.update(basis_single=bs, restart=10)      # several restart cycles at this size
        F = getattr(sim, "solve_total_field_" + pol)()
        its[bs] = sim.solver_info[pol]["iterations"]
        assert sim.solver_info[pol]["relres"] <= 1e-8
    ref = so.solve_total_field(so.cylinder_problem(N, eps_r=-1e6, pml_width=20, mask=40), pol)
    assert np.linalg.norm(F - ref) <= 1e-6 * np.linalg.norm(ref)
    assert its[1] <= 1.15 * its[0] + 3


@pytest.mark.parametrize("pol", ["TE", "TM"])
@pytest.mark.parametrize("Nx,Ny,er,angle", [(260, 180, 4.0 - 0.2j, 20.0), (150, 310, -1e6, 70.0)])
def test_scattering_rectangular_grids_default_options(Nx, Ny, er, angle, pol):
    """Non-square grids (dx != dy, more rows than columns and vice versa), lossy dielectric and metal-like
    scatterers, oblique incidence: the drop-in class with its DEFAULT options (polarisation-dependent restart /
    shift, residual tolerance derived from the 1e-6 field bar) against the oracle's direct solve."""
    import fdfd_b200
    f0 = 10e9
    lam = 299792458 / f0
    Lx, Ly = Nx / 45 * lam, Ny / 55 * lam
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, Lx, Ly, Nx, Ny)
    p = so.make_problem(f0, Lx, Ly, Nx, Ny)
    inside = ((sim.X / (0.9 * lam)) ** 2 + (sim.Y / (0.5 * lam)) ** 2) <= 1.0
    sim.add_object(er, 1.0, inside)
    so.add_object(p, er, 1.0, np.broadcast_to(inside, (Ny, Nx)))
    sim.add_UPML(pml_width=30, sigma_max=8)
    so.add_upml(p, pml_width=30, sigma_max=8)
    sim.add_mask(55)
    so.add_mask(p, 55)
    sim.add_source("plane_wave", angle_deg=angle, polarization="TE")
    so.add_source(p, "plane_wave", angle_deg=angle, polarization="TE")
    F = getattr(sim, "solve_total_field_" + pol)()
    ref = so.solve_total_field(p, pol)
    info = sim.solver_info[pol]
    assert info["converged"] and info["relres"] <= info["rtol"] <= 1e-8
    assert F.shape == (Ny, Nx)
    assert np.linalg.norm(F - ref.reshape(Ny, Nx)) <= 1e-6 * np.linalg.norm(ref)
