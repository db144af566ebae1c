"""GPU parity of K2 (matrix-free Helmholtz / curl-curl stencil) against the oracle: the numpy
matrix-free restatement for Dirichlet walls and explicit sparse products of the oracle's
yeeder2d operators for Bloch boundaries.  Tolerances: complex128 1e-13 relative L2 (pure
rounding: the kernel sums the same <= 9 products in a different order), complex64 2e-6."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import scattering_oracle as so, yee_oracle

pytestmark = pytest.mark.gpu

TOL = {"complex128": 1e-13, "complex64": 2e-6}


def _rand(rng, shape, lossy=True):
    a = rng.uniform(0.5, 2.0, shape) + 1j * rng.uniform(-1.0, 1.0, shape) * (1 if lossy else 0)
    return a


def _explicit(pol, Nx, Ny, ca, cb, cd, sx, sy, bc, kinc):
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([Nx, Ny], [1.0, 1.0], list(bc), kinc)
    a, b = sp.diags(ca.ravel()), sp.diags(cb.ravel())
    d = sp.diags(cd.ravel()) if cd is not None else 0 * sp.identity(Nx * Ny)
    if pol == "TE":
        return (sx * (DHX @ a @ DEX) + sy * (DHY @ b @ DEY) + d).tocsr()
    return (sx * (DEX @ a @ DHX) + sy * (DEY @ b @ DHY) + d).tocsr()


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("pol", ["TE", "TM"])
@pytest.mark.parametrize("shape", [(2, 2), (3, 5), (31, 17), (128, 64), (257, 130), (300, 300)])
def test_apply_dirichlet_vs_oracle(dtype, pol, shape):
    import torch
    import fdfd_b200
    Nx, Ny = shape
    rng = np.random.default_rng(Nx * 1000 + Ny)
    ca, cb, cd = _rand(rng, (Ny, Nx)), _rand(rng, (Ny, Nx)), _rand(rng, (Ny, Nx))
    sx, sy = 63.3, 41.7
    x = rng.standard_normal((Ny, Nx)) + 1j * rng.standard_normal((Ny, Nx))
    ref = so.apply_stencil(pol, ca, cb, cd, sx, sy, x)
    A = fdfd_b200.HelmholtzOperator2D(pol, Nx, Ny, ca, cb, cd, sx, sy, dtype=dtype)
    for variant in range(9):
        A.set_variant(variant)
        y = A.apply(torch.from_numpy(x).cuda().to(A.tdtype)).cpu().numpy().reshape(Ny, Nx)
        err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
        assert err <= TOL[dtype], (variant, err)
    yh = A.apply_host(x.ravel()).reshape(Ny, Nx)     # host-buffer C-ABI entry point
    assert np.linalg.norm(yh - ref) / np.linalg.norm(ref) <= TOL[dtype]


@pytest.mark.parametrize("pol", ["TE", "TM"])
@pytest.mark.parametrize("bc", [(1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("shape", [(2, 3), (16, 16), (40, 40), (65, 33)])
def test_apply_bloch_vs_explicit(pol, bc, shape):
    import torch
    import fdfd_b200
    Nx, Ny = shape
    rng = np.random.default_rng(7 + Nx)
    kinc = rng.standard_normal(2)
    ca, cb = _rand(rng, (Ny, Nx)), _rand(rng, (Ny, Nx))
    cd = _rand(rng, (Ny, Nx))
    sx, sy = -1.7, -2.9            # band-solver sign convention (Band_Diagram_Solver.py:312,319)
    phase = (np.exp(-1j * kinc[0] * Nx * 1.0), np.exp(-1j * kinc[1] * Ny * 1.0))
    M = _explicit(pol, Nx, Ny, ca, cb, cd, sx, sy, bc, kinc)
    x = rng.standard_normal(Nx * Ny) + 1j * rng.standard_normal(Nx * Ny)
    ref = M @ x
    A = fdfd_b200.HelmholtzOperator2D(pol, Nx, Ny, ca, cb, cd, sx, sy, bc=bc, phase=phase)
    y = A.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 1e-13
    # no diagonal term (cd = None), as in the band solver's A
    M0 = _explicit(pol, Nx, Ny, ca, cb, None, sx, sy, bc, kinc)
    A0 = fdfd_b200.HelmholtzOperator2D(pol, Nx, Ny, ca, cb, None, sx, sy, bc=bc, phase=phase)
    y0 = A0.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.linalg.norm(y0 - M0 @ x) / np.linalg.norm(M0 @ x) <= 1e-13


def test_apply_golden_small(golden):
    """A@x of the unmodified reference (120x120 cylinder problem with UPML), both polarisations."""
    import torch
    import fdfd_b200
    g = golden("scatter_small.npz")
    N = int(g["N"])
    p = so.cylinder_problem(N, eps_r=float(g["eps"]), pml_width=int(g["pml"]), mask=int(g["mask"]))
    for pol in ("TE", "TM"):
        ca, cb, cd, sx, sy = so.stencil_coefficients(p, pol)
        A = fdfd_b200.HelmholtzOperator2D(pol, N, N, ca, cb, cd, sx, sy)
        y = A.apply(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
        ref = g[f"y_{pol}"]
        assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 1e-13


def test_apply_linearity_full_size():
    """4096^2 (BASELINE metric size): size-independent properties — linearity and agreement of the
    two polarisation types on symmetric coefficients — plus a spot check of 4 rows vs the oracle."""
    import torch
    import fdfd_b200
    N = 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    def rnd():
        return torch.complex(torch.rand(N, N, generator=g, device="cuda", dtype=torch.float64) + 0.5,
                             torch.rand(N, N, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    ca, cb, cd = rnd(), rnd(), rnd()
    x1, x2 = rnd(), rnd()
    A = fdfd_b200.HelmholtzOperator2D("TM", N, N, ca, cb, cd, 63.3, 63.3)
    alpha = 0.3 - 1.7j
    lhs = A.apply(x1 + alpha * x2)
    rhs = A.apply(x1) + alpha * A.apply(x2)
    assert (torch.linalg.vector_norm(lhs - rhs) / torch.linalg.vector_norm(rhs)).item() <= 1e-14
    # host-buffer entry point at full size: pipelined in row chunks (upload / compute / download
    # overlap) -- must equal the device-resident apply bit for bit, for pinned and pageable buffers
    ref_dev = A.apply(x1).cpu()
    xh = torch.empty(N * N, dtype=torch.complex128, pin_memory=True)
    yh = torch.empty(N * N, dtype=torch.complex128, pin_memory=True)
    xh.copy_(x1.reshape(-1))
    A.apply_host(xh.numpy(), out=yh.numpy())
    assert torch.equal(yh, ref_dev.reshape(-1))
    ypage = A.apply_host(x1.reshape(-1).cpu().numpy())
    assert np.array_equal(ypage, ref_dev.reshape(-1).numpy())
    # first and last 4 rows against the numpy oracle evaluated on a 5-row band (the band's
    # artificial wall row is excluded; the true walls at j=0 / j=N-1 are exercised)
    full = A.apply(x1).reshape(N, N)
    for band, keep in ((slice(0, 5), slice(0, 4)), (slice(N - 5, N), slice(1, 5))):
        sub = so.apply_stencil("TM", ca[band].cpu().numpy(), cb[band].cpu().numpy(), cd[band].cpu().numpy(),
                               63.3, 63.3, x1[band].cpu().numpy())
        got = full[band].cpu().numpy()
        err = np.linalg.norm(got[keep] - sub[keep]) / np.linalg.norm(sub[keep])
        assert err <= 1e-13


@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_apply_4096_full_grid_vs_oracle(pol):
    """the BASELINE-size operator (4096^2 scaled cylinder problem with UPML, complex128)
    against the numpy restatement of A @ x on the WHOLE grid, <= 1e-13 relative (the oracle's explicit sparse
    A equals its matrix-free form to 1e-16, tests/test_oracle_golden.py)."""
    import torch
    import fdfd_b200
    N = 4096
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=40, sigma_max=10)
    ca, cb, cd, sx, sy = sim.stencil_coefficients(pol)
    rng = np.random.default_rng(0)                       # SURVEY §8d apply micro-benchmark vector
    x = rng.standard_normal(N * N) + 1j * rng.standard_normal(N * N)
    y = sim.operator(pol).apply(torch.from_numpy(x).cuda()).cpu().numpy()
    sim.close()
    ref = so.apply_stencil(pol, ca, cb, cd, sx, sy, x).ravel()
    err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    assert err <= 1e-13, err
    # and row by row, so a defect confined to a wall / PML row cannot hide in the global norm
    d = np.abs(y - ref).reshape(N, N).max(axis=1) / np.abs(ref).reshape(N, N).max(axis=1)
    assert d.max() <= 1e-12, (int(d.argmax()), d.max())


def test_apply_rejects_bad_input():
    import torch
    import fdfd_b200
    one = np.ones((8, 8))
    A = fdfd_b200.HelmholtzOperator2D("TE", 8, 8, one, one, one, 1.0, 1.0)
    with pytest.raises(ValueError):
        A.apply(torch.zeros(63, dtype=torch.complex128, device="cuda"))
    with pytest.raises(TypeError):
        A.apply(np.zeros(64, complex))
    with pytest.raises(ValueError):
        fdfd_b200.HelmholtzOperator2D("XX", 8, 8, one, one, one, 1.0, 1.0)
    with pytest.raises(ValueError):
        fdfd_b200.HelmholtzOperator2D("TE", 1, 8, np.ones((8, 1)), np.ones((8, 1)), None, 1.0, 1.0)
