"""CPU tests of the drop-in boundary: the shared library builds, loads and exports every
symbol include/fdfd_b200.h declares; argument validation works without a GPU; and compute
entry points refuse to run (no CPU fallback) when no device is visible."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from fdfd_b200 import build, _lib
    build.build()
    return _lib.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fdfd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdfd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from fdfd_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fdfd_b200.h but not exported"
    # and the ctypes table covers the header
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)


def test_opts_struct_layout(lib):
    from fdfd_b200 import _lib
    o = _lib.default_opts()
    assert o.rtol == 1e-8 and o.restart == 30 and o.mg_shift_im == 0.4 and o.reorth == 0 and o.mg_graph == 1
    assert o.mg_single == 0 and o.basis_single == 0 and o.mg_fuse_len == 1024 and o.mg_kh_max == 1.3
    # same size and the same field names in the same order as the C struct of the header
    assert C.sizeof(_lib.KrylovOpts) == lib.fdfd_krylov_opts_size()
    import re
    hdr = open(os.path.join(ROOT, "include", "fdfd_b200.h")).read()
    end = hdr.index("} fdfd_krylov_opts;")
    body = hdr[hdr.rindex("typedef struct", 0, end):end]
    body = body[body.index("{") + 1:]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [name.strip() for d in body.split(";") if d.strip()
              for name in d.strip().split(None, 1)[1].split(",")]
    assert fields == [f[0] for f in _lib.KrylovOpts._fields_]


def test_nnz_and_argument_validation(lib):
    from fdfd_b200 import _lib
    nx, ny, cx, cy = C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
    assert lib.fdfd_yee2d_csr_nnz(300, 300, 0, 0, C.byref(nx), C.byref(ny), C.byref(cx), C.byref(cy)) == 0
    assert (nx.value, ny.value, cx.value, cy.value) == (179700, 179700, 0, 0)
    assert lib.fdfd_yee2d_csr_nnz(4, 3, 1, 0, C.byref(nx), C.byref(ny), C.byref(cx), C.byref(cy)) == 0
    assert (nx.value, ny.value, cx.value, cy.value) == (24, 20, 1, 0)
    assert lib.fdfd_yee2d_csr_nnz(1, 5, 0, 0, C.byref(nx), C.byref(ny), C.byref(cx), C.byref(cy)) == 0
    assert (nx.value, cx.value) == (5, 1)
    assert lib.fdfd_yee2d_csr_nnz(0, 3, 0, 0, None, None, None, None) == _lib.FDFD_ERR_ARG
    assert lib.fdfd_yee2d_csr_nnz(3, 3, 2, 0, None, None, None, None) == _lib.FDFD_ERR_ARG
    assert b"BC" in lib.fdfd_last_error()


def test_no_cpu_fallback(lib):
    import fdfd_b200
    from fdfd_b200 import _lib
    if lib.fdfd_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(fdfd_b200.FdfdError):
        fdfd_b200.yeeder2d([4, 3], [0.5, 0.25], [0, 0])
    with pytest.raises(fdfd_b200.FdfdError):
        fdfd_b200.HelmholtzOperator2D("TE", 8, 8, np.ones((8, 8)), np.ones((8, 8)), None, 1.0, 1.0)
    sim = fdfd_b200.FDFD2DScatteringSolver(10e9, 0.06, 0.06, 32, 32)
    sim.add_source("plane_wave")
    with pytest.raises(fdfd_b200.FdfdError):
        sim.solve_total_field_TE()
    # the newer entry points refuse as well (argument checks first, then the device check), and the memory cache is
    # empty on a box that never allocated
    buf = (C.c_double * 8)()
    assert lib.fdfd_tsqr_r(2, 2, buf, None, 2, 0.0, 0.0, buf, None) == _lib.FDFD_ERR_CUDA
    assert lib.fdfd_tsqr_r(65, 100, buf, None, 100, 0.0, 0.0, buf, None) == _lib.FDFD_ERR_ARG
    assert lib.fdfd_tsqr_r(2, 4, buf, None, 3, 0.0, 0.0, buf, None) == _lib.FDFD_ERR_ARG
    assert fdfd_b200.trim_memory() == 0
    with pytest.raises(fdfd_b200.FdfdError):
        fdfd_b200.ModeSolver2D(50e9, 2e-3, 2e-3, 8, 8, 1).solve()


def test_product_does_not_import_oracle():
    """The package must never route through oracle/ (that would void every parity claim)."""
    pkg = os.path.join(ROOT, "fdfd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "ref_loader" not in src, f


def test_host_setup_matches_oracle():
    """Host-side problem definition of the drop-in class == oracle (== reference) bit for bit."""
    import fdfd_b200
    from oracle import scattering_oracle as so
    N = 96
    f0 = 10e9
    lam = 299792458 / f0
    sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
    sim.add_object(4.0, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
    sim.add_UPML(pml_width=16, sigma_max=10)
    sim.add_mask(24)
    sim.add_source("plane_wave", angle_deg=45.0)
    p = so.cylinder_problem(N, eps_r=4.0, pml_width=16, mask=24)
    for k in p.mats:
        assert np.array_equal(p.mats[k], getattr(sim, k)), k
    assert np.array_equal(p.source, sim.source)
    assert np.array_equal(p.q, sim.Q.diagonal())
    ca, cb, cd, sx, sy = sim.stencil_coefficients("TM")
    oa, ob, od, osx, osy = so.stencil_coefficients(p, "TM")
    assert np.array_equal(ca, oa) and np.array_equal(cb, ob) and np.array_equal(cd, od) and (sx, sy) == (osx, osy)
    with pytest.raises(ValueError):
        sim.add_source("point")
    with pytest.raises(ValueError):
        sim.add_mask(np.ones((3, 3)))
    with pytest.raises(TypeError):
        sim.add_mask({})


def test_amplification_model_covers_the_measurements(golden):
    """default_rtol's model 71 sqrt(L / lambda) is an upper envelope of the factors measured with the
    reference's LU at 300^2, 1024^2 and 2048^2 (both polarisations)."""
    import fdfd_b200
    g = golden("scatter_2048.npz")
    f0 = 10e9
    lam = 299792458 / f0
    for N in (300, 1024, 2048):
        sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N)
        model = 1e-6 / (1.25 * sim.default_rtol(1e-6)) if sim.default_rtol(1e-6) < 1e-8 else None
        for pol in ("TM", "TE"):
            amp = float(g[f"amp_{N}_{pol}"])
            assert amp <= sim.AMPLIFICATION_PER_SQRT_WAVELENGTH * np.sqrt(N / 50)
            if model is not None:
                assert amp <= model * 1.0000001
