"""GPU parity of the full-vector mode solver (factored Omega = P Q on the device, K6 Krylov-Schur
shift-invert with the K3 inner solve) against golden results of the unmodified reference.
Bar (BASELINE.json): eigenvalues / n_eff within 1e-8 relative."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.mode_cases()


@pytest.mark.parametrize("case", ["strip", "shapes", "ridge"])
def test_mode_solver_vs_golden(golden, case):
    import fdfd_b200
    g = golden("mode_golden.npz")
    s = _cases()[case](fdfd_b200.ModeSolver2D)
    s.solve()
    ev, neff = g[f"{case}_eigenvalues"], g[f"{case}_neff"]
    assert s._resolve_eigs_guess(None) == float(np.real(g[f"{case}_sigma"]))
    assert np.max(np.abs(s.eigenvalues - ev) / np.abs(ev)) <= 1e-8
    assert np.max(np.abs(s.neff - neff) / np.abs(neff)) <= 1e-8
    assert np.allclose(s.propagation_constant, np.real(neff), rtol=1e-8, atol=0)
    assert s.eigenvectors.shape == (s.n_e, s.num_modes)
    for fld, shape in (("Ex", s.shape_ex), ("Ey", s.shape_ey), ("Ez", s.shape_ez), ("Hx", s.shape_hx),
                       ("Hy", s.shape_hy), ("Hz", s.shape_hz)):
        a = getattr(s, fld)
        assert a.shape == (*shape, s.num_modes)
        nrm = np.array([np.linalg.norm(a[:, :, m]) for m in range(s.num_modes)])
        ref_n = g[f"{case}_{fld}_norm"]
        assert np.allclose(nrm, ref_n, rtol=1e-6, atol=1e-9 * ref_n.max())
        dec, ref = a[::7, ::5, :], g[f"{case}_{fld}_dec"]
        for m in range(s.num_modes):
            scale = max(np.linalg.norm(ref[:, :, m]), 1e-9 * ref_n.max())
            if fld in ("Ex", "Ey"):
                # the reference rotates Ex/Ey twice (they alias `eigenvectors`), which leaves them with a
                # run-dependent phase: compare up to a free phase
                ph = np.vdot(ref[:, :, m], dec[:, :, m])
                ph = ph / abs(ph) if abs(ph) > 0 else 1.0
                d = np.linalg.norm(dec[:, :, m] - ph * ref[:, :, m])
            else:                          # the most-real phase leaves a sign free
                d = min(np.linalg.norm(dec[:, :, m] - ref[:, :, m]), np.linalg.norm(dec[:, :, m] + ref[:, :, m]))
            assert d <= 1e-6 * scale, (fld, m, d / scale)
    assert np.all(s.solver_info["residuals"] < 1e-8)


def test_mode_solver_errors():
    import fdfd_b200
    with pytest.raises(ValueError):
        fdfd_b200.ModeSolver2D(1e9, 1.0, 1.0, 0, 4, 1)
    s = fdfd_b200.ModeSolver2D(30e9, 4e-3, 4e-3, 4, 4, 3)
    s.add_pec((0, 4), (0, 4))
    with pytest.raises(ValueError):
        s.solve()
    with pytest.raises(ValueError):
        s.add_rectangle(2.0, 1.0, (0, 9), (0, 2))
    with pytest.raises(ValueError):
        s.add_pml(direction="z")
