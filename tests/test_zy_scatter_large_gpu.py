"""(Runs late in the suite.)  Parity of the drop-in scattering solver at the sizes whose reference (SuperLU)
solution fits the build container — SURVEY.md §8c "scaled 1024^2" and "scaled 2048^2" cases — against
tests/golden/scatter_1024.npz / scatter_2048.npz (made by tests/golden/make_golden.py from the unmodified
reference; 2048^2 costs it 228 s + 205 s and 17.6 GB).  Bars of BASELINE.json: residual <= 1e-8, field
relative L2 error <= 1e-6 (complex128) — with the solver's DEFAULT options: the residual tolerance is derived
from the field tolerance by FDFD2DScatteringSolver.default_rtol()."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,probe", [(1024, (512, 256)), (2048, (1024, 512))])
@pytest.mark.parametrize("pol", ["TM", "TE"])
def test_scattering_large_vs_reference(golden, N, probe, pol):
    import fdfd_b200
    g = golden(f"scatter_{N}.npz")
    assert int(g["N"]) == N
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(-1e6, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=40, sigma_max=10)
    sim.add_mask(80)
    sim.add_source("plane_wave", angle_deg=45.0, polarization="TE")
    assert sim.solver_options["rtol"] is None          # default options: nothing overridden here
    F = getattr(sim, "solve_total_field_" + pol)()
    sim.close()
    info = sim.solver_info[pol]
    assert info["rtol"] == sim.default_rtol(1e-6) and info["relres"] <= info["rtol"] <= 1e-8
    if N == 2048:
        # the bound the default tolerance is built on: measured amplification (reference LU) x residual
        assert float(g[f"amp_2048_{pol}"]) * info["rtol"] <= 1e-6
    dec = g[f"dec_{pol}"]
    assert np.linalg.norm(F[::32, ::32] - dec) <= 1e-6 * np.linalg.norm(dec)
    assert abs(F[probe] - g[f"probe_{pol}"]) <= 1e-6 * g[f"max_{pol}"]
    assert abs(np.linalg.norm(F) - g[f"norm_{pol}"]) <= 1e-6 * g[f"norm_{pol}"]
