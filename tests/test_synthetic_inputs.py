"""The device-built synthetic problem of bench.py / the sharded solves (fdfd_b200/synthetic.py, torch tensor
ops evaluated per y-slab) is the reference example scaled per SURVEY.md §8d: its slabs must equal the arrays
the drop-in class builds on the host (which are bit-identical to the reference's, test_oracle_vs_reference.py).
Runs the torch code on CPU tensors."""
import numpy as np
import pytest


@pytest.mark.parametrize("pol", ["TE", "TM"])
@pytest.mark.parametrize("rows", [(0, 300), (100, 175), (262, 300)])
def test_slab_problem_matches_host_class(pol, rows):
    import fdfd_b200
    from fdfd_b200.synthetic import slab_problem
    N = 300
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=40, sigma_max=10)
    sim.add_mask(80)
    sim.add_source("plane_wave", angle_deg=45.0, polarization="TE")
    ca, cb, cd, sx, sy = sim.stencil_coefficients(pol)
    j0, j1 = rows
    tca, tcb, tcd, tsx, tsy, src, q = slab_problem(N, N, j0, j1, pol=pol, device="cpu")
    assert (tsx, tsy) == (sx, sy)
    for got, ref in ((tca, ca[j0:j1]), (tcb, cb[j0:j1]), (tcd, cd[j0:j1]), (src, sim.source.reshape(N, N)[j0:j1])):
        assert np.linalg.norm(got.numpy() - ref) <= 1e-14 * np.linalg.norm(ref)
    assert np.array_equal(q.numpy().real, sim.Q.diagonal().reshape(N, N)[j0:j1])
