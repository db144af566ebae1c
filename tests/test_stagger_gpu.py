"""K1b on the GPU (csrc/yee_assemble.cu: stagger_fill_kernel, C-ABI fdfd_stagger_csr_fill*) against the oracle,
byte for byte: the rectangular staggered difference operators of ModeSolver2D (Mode_Solver_2D.py:482-544) and
the sixteen operators of PeriodicModeSolver3D incl. the periodic z difference / forward average
(Periodic_Solver_3D.py:95-194).  The oracle's versions are pinned against the unmodified reference in
tests/test_oracle_vs_reference.py."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import mode_oracle as mo
from oracle import periodic3d_oracle as po

pytestmark = pytest.mark.gpu


def _same(a, b, name=""):
    assert a.shape == b.shape and a.format == b.format, name
    assert a.indptr.dtype == b.indptr.dtype == np.int32 and a.data.dtype == b.data.dtype, name
    assert a.indptr.tobytes() == b.indptr.tobytes() and a.indices.tobytes() == b.indices.tobytes(), name
    assert a.data.tobytes() == b.data.tobytes(), name


def _golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(m)
    except Exception:           # the module imports the reference loader at the top; only the case builders are used
        pass
    return m


@pytest.mark.parametrize("Nx,Ny,dx,dy", [(5, 4, 0.3, 0.7), (1, 1, 0.5, 0.25), (240, 160, 0.10471975511965977, 0.10471975511965977),
                                         (33, 1, 0.2, 0.9), (2, 57, 1.7, 0.01)])
def test_mode_solver_operators_bitwise(Nx, Ny, dx, dy):
    import fdfd_b200
    s = fdfd_b200.ModeSolver2D(50e9, 1.0, 1.0, Nx, Ny, 1)
    s.dx_normalized, s.dy_normalized = dx, dy
    got = s._yeeder2d()
    D = mo.operators(Nx, Ny, dx, dy)
    for g, name in zip(got, ("d_ez_ex", "d_ez_ey", "d_ey_hz", "d_ex_hz", "d_hy_ez", "d_hx_ez", "d_hz_hx", "d_hz_hy")):
        _same(g, D[name], name)


@pytest.mark.parametrize("shape", [(8, 7, 6), (3, 2, 2), (30, 30, 32), (1, 1, 2), (5, 9, 3)])
def test_periodic3d_operators_bitwise(shape):
    import fdfd_b200
    Nx, Ny, Nz = shape
    s = fdfd_b200.PeriodicModeSolver3D(Nx=Nx, Ny=Ny, Nz=Nz, x_range=4e-3, y_range=3.5e-3, z_range=3e-3, freq=25e9,
                                       num_modes=1, sigma_guess=0)
    D = po.operators(Nx, Ny, Nz, s.dx, s.dy, s.dz)
    for name in s._OPERATORS:
        _same(getattr(s, name), D[name], name)


def test_device_output_and_validation():
    """Device-pointer entry point (operators left on the GPU) equals the host one; bad shapes are refused."""
    import ctypes as C
    import torch
    from fdfd_b200 import _lib, stagger
    L = _lib.lib()
    ins, outs = (9, 6, 4), (8, 6, 4)
    ref, refh = stagger.difference(ins, outs, 0.37, 0, with_h=True)
    rows = int(np.prod(outs))
    ptr = torch.empty(rows + 1, dtype=torch.int32, device="cuda")
    idx = torch.empty(2 * rows, dtype=torch.int32, device="cuda")
    dat = torch.empty(2 * rows, dtype=torch.float64, device="cuda")
    hd = torch.empty(2 * rows, dtype=torch.float64, device="cuda")
    _lib.check(L.fdfd_stagger_csr_fill((C.c_int * 3)(*ins), (C.c_int * 3)(*outs), 0, -1.0 / 0.37, 1.0 / 0.37, ptr.data_ptr(),
                                       idx.data_ptr(), dat.data_ptr(), hd.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(ptr.cpu().numpy(), ref.indptr) and np.array_equal(idx.cpu().numpy(), ref.indices)
    assert dat.cpu().numpy().tobytes() == ref.data.tobytes() and hd.cpu().numpy().tobytes() == refh.data.tobytes()
    with pytest.raises(ValueError):
        stagger.difference((8, 6, 4), (8, 6, 4), 1.0, 0)          # no forward neighbour in the last column
    with pytest.raises(ValueError):
        stagger.periodic_z((4, 4, 1), 0.5, 0.5)                    # a one-plane ring would need a merged entry


def test_factored_operator_equals_reference_omega(golden):
    """Omega = P_e Q_e built from the GPU-assembled operators equals the oracle's (reference-pinned) Omega, and the
    3-D pencil (A, B) equals the oracle's, for the fixture cases."""
    import fdfd_b200
    m = _golden_module()
    build = m.mode_cases()["strip"]
    s = build(fdfd_b200.ModeSolver2D)
    f = s._factors()
    cell = {f"{p}_{c}": getattr(s, f"cell_{p}_r_{c}") for p in ("eps", "mu") for c in ("xx", "yy", "zz")}
    pec = {c: getattr(s, f"pec_{c}_mask") for c in ("xx", "yy", "zz")}
    pmc = {c: getattr(s, f"pmc_{c}_mask") for c in ("xx", "yy", "zz")}
    o = mo.solve(cell, s.material_no_average_mask, pec, pmc, s.Nx, s.Ny, s.dx_normalized, s.dy_normalized,
                 s.num_modes, s._resolve_eigs_guess(None), s._has_lossy_material())
    Om = f["P_e"] @ f["Q_e"]
    assert abs(Om - o["Omega"]).max() <= 1e-12 * abs(o["Omega"]).max()
    s3, _ = m.periodic3d_cases()["tiny"](fdfd_b200.PeriodicModeSolver3D)
    A, B = s3._build_eigen_matrices()
    cell3 = {k: getattr(s3, f"cell_{k[0].upper()}{k[1:]}_3D") for k in ("erxx", "eryy", "erzz", "mrxx", "mryy", "mrzz")}
    D = po.operators(s3.Nx, s3.Ny, s3.Nz, s3.dx, s3.dy, s3.dz)
    Ao, Bo = po.pencil(po.materials_on_fields(cell3, s3.material_no_average_mask), D, s3.omega)
    assert abs(A - Ao).max() <= 1e-12 * abs(Ao).max() and abs(B - Bo).max() == 0
