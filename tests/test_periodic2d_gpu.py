"""GPU parity of the 2-D periodic mode solver (K1b operators, device CSR pencil, Krylov-Schur shift-invert with the
plane-block / dense shifted solve) against golden results of the unmodified reference.

The reference's ARPACK call converges only some of the pairs it returns (residuals of the rest: 1e-5 ... 1, and they
change from run to run — tests/golden/make_golden.py stores the residual of every pair): parity is asserted on the
converged pairs (eigenvalues 1e-8 relative, fields 1e-6), and every pair THIS solver returns must be an eigenpair
of the pencil (relative residual <= 1e-7, the matched ones <= 1e-9)."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.periodic2d_cases()


@pytest.mark.parametrize("shape", [(7, 5), (40, 12), (200, 80)])
def test_periodic2d_operators_bitwise(shape):
    """The twelve operators from the K1b kernel against the oracle's (which equal the reference's bit for bit,
    tests/test_oracle_vs_reference.py)."""
    import fdfd_b200
    from conftest import same_sparse
    from oracle import periodic2d_oracle as po
    Nx, Nz = shape
    s = fdfd_b200.PeriodicModeSolver2D("TM", 20e9, 1e-3 * Nx / 20, 0.7e-3 * Nz / 10, Nx, Nz, 2)
    D = po.operators(Nx, Nz, s.dx, s.dz)
    for name in s._OPERATORS:
        assert same_sparse(getattr(s, name), D[name]), name


@pytest.mark.parametrize("case", ["grating_tm", "slab_te", "antenna"])
def test_periodic2d_vs_golden(golden, case):
    import fdfd_b200
    from oracle import periodic2d_oracle as po
    g = golden("periodic2d_golden.npz")
    s = _cases()[case](fdfd_b200.PeriodicModeSolver2D)
    s.solve()
    assert s.eigenvalues.shape == (s.num_modes,) and s.eigenvectors.shape[1] == s.num_modes
    assert np.all(np.diff(np.abs(s.eigenvalues - s.guess)) >= -1e-9 * np.abs(s.eigenvalues).max())   # ordered by |lambda - guess|
    assert np.allclose(s.neff, s.eigenvalues / s.k0) and np.allclose(s.propagation_constant, np.imag(s.neff))
    # every returned pair is an eigenpair of the pencil the reference builds (checked on the host with scipy)
    A, B, free = s._pencil()
    res = po.relative_residuals(A, B, s.eigenvalues, s.eigenvectors[free, :])
    # (the class equilibrates the pencil first: with the 1e8 penalty materials A holds entries of 1e-6 ... 1e13 and
    # even a dense LAPACK eigen-solve of the raw pencil stops at 1e-9 in this measure; the reference's unconverged
    # pairs sit at 1e-5 ... 1)
    assert np.max(res) <= 1e-7, res
    # the class reports the same measure, evaluated on the device through the equilibrated operators: at 1e-11 both
    # are rounding-level quantities and agree in magnitude only
    assert np.allclose(res, s.solver_info["residuals"], rtol=0.5, atol=1e-10)
    first, second = ("Ex", "Hy") if s.polarization == "TM" else ("Hx", "Ey")
    conv = g[f"{case}_converged"]
    assert conv.size >= 1
    for pos, k in enumerate(conv):
        ref = g[f"{case}_eigenvalues"][k]
        j = int(np.argmin(np.abs(s.eigenvalues - ref)))
        assert abs(s.eigenvalues[j] - ref) <= 1e-8 * abs(ref), (case, k, s.eigenvalues, ref)
        assert res[j] <= 1e-9, res
        # fields: the pair (first, second) shares one arbitrary complex factor
        mine = np.concatenate([getattr(s, f)[:, j].reshape(getattr(s, "shape_" + f.lower()), order="F")[::3, ::2].ravel()
                               for f in (first, second)])
        gold = np.concatenate([g[f"{case}_{f}_dec"][pos].ravel() for f in (first, second)])
        c = np.vdot(gold, mine) / np.vdot(gold, gold)
        assert np.linalg.norm(mine - c * gold) <= 1e-6 * np.linalg.norm(mine)


def test_periodic2d_errors_and_invalidation():
    import fdfd_b200
    with pytest.raises(ValueError):
        fdfd_b200.PeriodicModeSolver2D("TEM", 20e9, 1e-3, 1e-3, 8, 8, 2)
    s = fdfd_b200.PeriodicModeSolver2D("TE", 20e9, 4e-3, 2e-3, 24, 8, 2, guess=300j)
    s.add_rectangle(4.0, 1, (0.5e-3, 1.5e-3), (0, 8))
    s.solve()
    assert s.Hx.shape == (s.n_hx, 2) and s.Ey.shape == (s.n_ey, 2)
    assert s._field_array(s.Ey, 0, s.shape_ey).shape == s.shape_ey
    s.add_pml(pml_width=6, direction="x")
    assert s.eigenvalues is None and not hasattr(s, "Ey")
    big = fdfd_b200.PeriodicModeSolver2D("TM", 20e9, 1e-3, 1e-3, 2, 2, 8)
    with pytest.raises(ValueError):
        big.solve()
