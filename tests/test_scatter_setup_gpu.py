"""Device set-up of FDFD2DScatteringSolver: coefficient arrays and right-hand side built on the
device from the recorded add_* calls — object-id table fill (csrc: scatter_fill_kernel) + O(perimeter) UPML frame
and source band evaluated by numpy — must be BIT-identical to the host path (the reference's full arrays,
Scattering_Solver_2D.py:72-188, which tests/test_oracle_vs_reference.py pins against the reference itself)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F0 = 10e9
LAM = 299792458 / F0


def _build(Nx, Ny, case, **kw):
    import fdfd_b200
    sim = fdfd_b200.FDFD2DScatteringSolver(F0, Nx / 50 * LAM, Ny / 50 * LAM, Nx, Ny, **kw)
    X, Y = sim.X, sim.Y
    if case == "cylinder":
        sim.add_object(-1e6, 1.0, (X ** 2 + Y ** 2) <= (0.5 * LAM) ** 2)
        sim.add_UPML(pml_width=min(40, Nx // 6), sigma_max=10)
        sim.add_mask(min(80, Nx // 4))
    elif case == "two_objects_into_pml":
        sim.add_object((4.0, 3.0 + 0.2j, 2.5), (1.0, 1.5, 1.2 - 0.1j), np.abs(Y) < 0.2 * LAM)        # slab through the x-PML
        sim.add_object(9.0 + 1j, 1.0, (np.abs(X - 0.3 * LAM) < 0.25 * LAM) & (np.abs(Y) < 0.3 * LAM))   # overrides part of it
        sim.add_UPML(pml_width=12, n=2, sigma_max=7.5)
        sim.add_mask(20)
    elif case == "x_then_y_pml":
        sim.add_object(2.2, 1.0, (X ** 2 + (Y - 0.1 * LAM) ** 2) <= (0.4 * LAM) ** 2)
        sim.add_UPML(pml_width=10, sigma_max=5, direction="x")
        sim.add_UPML(pml_width=16, sigma_max=9, direction="y")
        sim.add_mask(24)
    elif case == "no_mask":
        sim.add_object(3.0, 1.0, np.abs(X) < 0.2 * LAM)
        sim.add_UPML(pml_width=8, sigma_max=4)
    sim.add_source("plane_wave", angle_deg=33.0, polarization="TE", amplitude=1.3)
    return sim


@pytest.mark.parametrize("Nx,Ny,case", [(120, 120, "cylinder"), (300, 300, "cylinder"), (200, 150, "two_objects_into_pml"),
                                        (96, 160, "x_then_y_pml"), (64, 64, "no_mask")])
@pytest.mark.parametrize("pol", ["TE", "TM"])
def test_device_setup_bit_identical(Nx, Ny, case, pol):
    import torch
    dev_sim = _build(Nx, Ny, case)
    host_sim = _build(Nx, Ny, case)
    host_sim.ERzz                                   # materialise: the host path takes over
    assert dev_sim._host is None and host_sim._host is not None
    A = dev_sim.operator(pol)
    assert dev_sim._host is None                    # nothing materialised the (Ny, Nx) host arrays
    ca, cb, cd, sx, sy = host_sim.stencil_coefficients(pol)
    for got, ref, name in ((A.ca, ca, "ca"), (A.cb, cb, "cb"), (A.cd, cd, "cd")):
        assert got.cpu().numpy().tobytes() == np.ascontiguousarray(ref).tobytes(), name
    assert (A.sx, A.sy) == (sx, sy)
    b_dev = dev_sim._rhs(A)
    Ah = host_sim.operator(pol)
    b_host = host_sim._rhs(Ah)
    assert torch.equal(b_dev, b_host)
    assert dev_sim._host is None
    if case != "no_mask":
        assert torch.linalg.vector_norm(b_dev).item() > 0
    dev_sim.close(); host_sim.close()


def test_device_setup_slab_rows_and_solution():
    """Rows of a y-slab only (what a sharded rank builds), and the drop-in solve through the device set-up equals
    the solve through the host arrays bit for bit."""
    import torch
    N = 128
    full = _build(N, N, "cylinder")
    full.ERzz
    ca, cb, cd, _, _ = full.stencil_coefficients("TE")
    part = _build(N, N, "cylinder")
    part._rows = (40, 104)
    got = part._device_coefficients("TE")
    for g, r in zip(got[:3], (ca, cb, cd)):
        assert g.cpu().numpy().tobytes() == np.ascontiguousarray(r[40:104]).tobytes()
    a, b = _build(N, N, "cylinder"), _build(N, N, "cylinder")
    b.MRxx
    Fa, Fb = a.solve_total_field_TM(), b.solve_total_field_TM()
    assert a._host is None and np.array_equal(Fa, Fb)
    a.close(); b.close(); full.close()


def test_unrecordable_calls_fall_back_to_the_host_arrays():
    """Point sources, array masks and objects added after the UPML are outside the recorded recipe: the class
    materialises the reference's arrays and behaves as before."""
    import fdfd_b200
    N = 96
    sim = fdfd_b200.FDFD2DScatteringSolver(F0, N / 50 * LAM, N / 50 * LAM, N, N)
    sim.add_UPML(pml_width=10, sigma_max=5)
    sim.add_object(4.0, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.3 * LAM) ** 2)        # after the UPML -> host path
    assert sim._host is not None and sim.ERzz[N // 2, N // 2] == 4.0
    sim2 = fdfd_b200.FDFD2DScatteringSolver(F0, N / 50 * LAM, N / 50 * LAM, N, N)
    sim2.add_source("point", location=(0.0, 0.0))
    assert sim2._host is not None and np.count_nonzero(sim2.source) == N * N
    with pytest.raises(ValueError):
        sim2.add_source("point")
    with pytest.raises(ValueError):
        sim2.add_source("laser")
    q = np.ones((N, N)); q[20:-20, 20:-20] = 0
    sim3 = _build(N, N, "no_mask")
    sim3.add_mask(q)
    F = sim3.solve_total_field_TE()
    assert F.shape == (N, N) and np.isfinite(F).all()
    sim3.close()
