"""CPU tests: the oracle against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py) and the known-answer values of SURVEY.md §8c."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import scattering_oracle as so, yee_oracle


def _mat(g, n, name, M):
    cls = sp.csr_matrix if str(g[f"c{n}_{name}_format"]) == "csr" else sp.csc_matrix
    return cls((g[f"c{n}_{name}_data"], g[f"c{n}_{name}_indices"], g[f"c{n}_{name}_indptr"]), shape=(M, M))


def test_yee_golden_bitwise(golden):
    from conftest import same_sparse
    g = golden("yee2d_kat.npz")
    cases = g["cases"]
    for n, c in enumerate(cases):
        Nx, Ny, dx, dy, bx, by, kx, ky = c
        Nx, Ny, bx, by = int(Nx), int(Ny), int(bx), int(by)
        kinc = None if np.isnan(kx) else [kx, ky]
        ops = yee_oracle.yeeder2d([Nx, Ny], [dx, dy], [bx, by], kinc)
        for name, o in zip(("DEX", "DEY", "DHX", "DHY"), ops):
            assert same_sparse(_mat(g, n, name, Nx * Ny), o), (n, name)


def test_yee_survey_kat():
    # yeeder2d([4,3],[0.5,0.25],[0,0]) and the Bloch variant
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([4, 3], [0.5, 0.25], [0, 0])
    assert DEX.indptr.tolist() == [0, 2, 4, 6, 7, 9, 11, 13, 14, 16, 18, 20, 21]
    assert DEX.indices.tolist() == [0, 1, 1, 2, 2, 3, 3, 4, 5, 5, 6, 6, 7, 7, 8, 9, 9, 10, 10, 11, 11]
    assert set(DEX.data.tolist()) == {-2.0, 2.0} and DEX.dtype == np.float64
    assert DEY.indptr.tolist()[-5:] == [16, 17, 18, 19, 20]
    assert set(DEY.data.tolist()) == {-4.0, 4.0}
    assert DHX.format == "csc" and np.array_equal(DHX.data, -DEX.data)
    DEX, DEY, _, _ = yee_oracle.yeeder2d([4, 3], [0.5, 0.25], [1, 1], [0.3, 0.7])
    assert DEX.nnz == 24 and DEY.nnz == 24 and DEX.dtype == np.complex128
    assert DEX[3, 0] == 1.6506712298193567 - 1.1292849467900707j
    assert DEY[8, 0] == 3.4612957664917654 - 2.0048520186951913j
    assert DEX.indptr.dtype == np.int32 and DEX.indices.dtype == np.int32


def test_scatter_small_golden(golden):
    g = golden("scatter_small.npz")
    N = int(g["N"])
    p = so.cylinder_problem(N, eps_r=float(g["eps"]), pml_width=int(g["pml"]), mask=int(g["mask"]))
    for pol in ("TE", "TM"):
        A = so.build_A(p, pol)
        assert np.array_equal(A @ g["x"], g[f"y_{pol}"])
        assert np.array_equal(so.rhs(A, p), g[f"b_{pol}"])
        y = so.apply_stencil(pol, *so.stencil_coefficients(p, pol), g["x"]).ravel()
        assert np.linalg.norm(y - g[f"y_{pol}"]) <= 1e-14 * np.linalg.norm(g[f"y_{pol}"])
        F = so.solve_total_field(p, pol, A)
        assert np.linalg.norm(F - g[f"F_{pol}"]) <= 1e-12 * np.linalg.norm(F)


def test_scatter_cfg1_golden(golden):
    """BASELINE config 1 (300x300, eps=-1e6 cylinder): SURVEY.md §8c values."""
    g = golden("scatter_cfg1.npz")
    p = so.cylinder_problem(300)
    Hz = so.solve_total_field(p, "TM")
    assert abs(Hz[150, 75] - (0.160259857806 + 0.166863782383j)) < 1e-10
    assert abs(Hz[150, 75] - g["probe_TM"]) < 1e-12
    assert abs(np.abs(Hz).max() - 1.9723) < 1e-3
    assert np.linalg.norm(Hz[::10, ::10] - g["dec_TM"]) <= 1e-11 * np.linalg.norm(g["dec_TM"])


def test_scatter_1024_golden(golden):
    """SURVEY.md §8c scaled 1024^2 case: the oracle's direct solve against the reference's field (TM)."""
    g = golden("scatter_1024.npz")
    p = so.cylinder_problem(1024)
    Hz = so.solve_total_field(p, "TM")
    assert abs(Hz[512, 256] - (-0.846661030394 - 0.831841310624j)) < 1e-10        # value recorded in the survey
    assert abs(Hz[512, 256] - g["probe_TM"]) < 1e-11
    assert np.linalg.norm(Hz[::32, ::32] - g["dec_TM"]) <= 1e-10 * np.linalg.norm(g["dec_TM"])


def test_band_golden(golden):
    """Band oracle vs golden eigenvalues of the reference (config 3 at six path points)."""
    from oracle import band_oracle as bo
    g = golden("band_golden.npz")
    p = bo.BandProblem(1.0, 40, background_er=10.2)
    p.add_circle(0.4, er=1.0)
    bp, ticks = bo.bloch_path(bo.gamma_x_m_gamma(1.0, 1.0), 200)
    assert np.array_equal(bp, g["sq_path"]) and ticks == [0, 59, 117, 199]
    o = bo.band_structure(p, bp[:, g["sq_idx"][[1, 3]]], 5)
    for pol in ("TM", "TE"):
        ref = g[f"sq_ev_{pol}"][:, [1, 3]]
        assert np.max(np.abs(o[pol][0] - ref)) <= 1e-9 * np.abs(ref).max()
    assert np.allclose(o["TM"][1][:, 0], [0.1161753797, 0.2956241914, 0.3903181775, 0.3991116489, 0.4862472813], atol=1e-9)


@pytest.mark.parametrize("case", ["antenna", "slab_te", "grating_tm"])
def test_periodic2d_oracle_golden(golden, case):
    """oracle/periodic2d_oracle.py (operators, averaging, pencil, eigs call) on the cell tensors of the drop-in
    class reproduces the pairs the reference converged (relative residual <= 1e-9 in its own pencil)."""
    import importlib.util
    import os
    import fdfd_b200
    from oracle import periodic2d_oracle as po
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    s = m.periodic2d_cases()[case](fdfd_b200.PeriodicModeSolver2D)
    g = golden("periodic2d_golden.npz")
    cell = {f"{p}_{c}": getattr(s, f"cell_{p}_r_{c}") for p in ("eps", "mu") for c in ("xx", "yy", "zz")}
    mats = po.materials_on_fields(cell, s.material_no_average_mask)
    lam, X, res = po.solve(s.polarization, mats, s.Nx, s.Nz, s.dx, s.dz, s.omega, s.num_modes, s.guess)
    conv = g[f"{case}_converged"]
    assert conv.size >= 1
    for k in conv:
        ref = g[f"{case}_eigenvalues"][k]
        j = int(np.argmin(np.abs(lam - ref)))
        assert abs(lam[j] - ref) <= 1e-8 * abs(ref) and res[j] <= 1e-9
