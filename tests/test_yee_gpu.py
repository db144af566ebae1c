"""GPU parity of K1 (yeeder2d assembly kernel) — bit-exact against the oracle and the golden
fixtures, through the C-ABI (fdfd_yee2d_csr_fill_host / fdfd_yee2d_csr_fill)."""
import itertools

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import same_sparse
from oracle import yee_oracle

pytestmark = pytest.mark.gpu


def test_yeeder2d_bitwise_grid():
    import fdfd_b200
    rng = np.random.default_rng(1)
    for Nx, Ny in itertools.product([1, 2, 3, 4, 7, 16, 33], repeat=2):
        for bc in itertools.product([0, 1], repeat=2):
            for kinc in (None, [0.3, 0.7], rng.standard_normal(2)):
                res = [0.5, 0.25] if kinc is None else list(rng.uniform(0.01, 1, 2))
                G = fdfd_b200.yeeder2d([Nx, Ny], res, list(bc), kinc)
                O = yee_oracle.yeeder2d([Nx, Ny], res, list(bc), kinc)
                for g, o in zip(G, O):
                    assert same_sparse(g, o), (Nx, Ny, bc, kinc)


def test_yeeder2d_golden(golden):
    import fdfd_b200
    g = golden("yee2d_kat.npz")
    for n, c in enumerate(g["cases"]):
        Nx, Ny, dx, dy, bx, by, kx, ky = c
        Nx, Ny, bx, by = int(Nx), int(Ny), int(bx), int(by)
        kinc = None if np.isnan(kx) else [kx, ky]
        ops = fdfd_b200.yeeder2d([Nx, Ny], [dx, dy], [bx, by], kinc)
        for name, o in zip(("DEX", "DEY", "DHX", "DHY"), ops):
            cls = sp.csr_matrix if str(g[f"c{n}_{name}_format"]) == "csr" else sp.csc_matrix
            ref = cls((g[f"c{n}_{name}_data"], g[f"c{n}_{name}_indices"], g[f"c{n}_{name}_indptr"]),
                      shape=(Nx * Ny, Nx * Ny))
            assert same_sparse(ref, o), (n, name)


@pytest.mark.parametrize("N", [300, 1000])
def test_yeeder2d_large_and_device(N):
    """cfg1 size (M=90 000, nnz 179 700) and a non-power-of-two larger grid; device-resident variant."""
    import fdfd_b200
    res = [0.1256, 0.1256]
    for bc, kinc in (([0, 0], None), ([1, 1], [0.7, -1.3])):
        G = fdfd_b200.yeeder2d([N, N - 3], res, bc, kinc)
        O = yee_oracle.yeeder2d([N, N - 3], res, bc, kinc)
        for g, o in zip(G, O):
            assert same_sparse(g, o)
        D = fdfd_b200.yeeder2d_device([N, N - 3], res, bc, kinc)
        (px, ix, vx), (py, iy, vy) = yee_oracle.yeeder2d_arrays([N, N - 3], res, bc, kinc)
        assert np.array_equal(D["x"]["indptr"].cpu().numpy(), px)
        assert np.array_equal(D["x"]["indices"].cpu().numpy(), ix)
        assert D["x"]["e_data"].cpu().numpy().tobytes() == vx.tobytes()
        assert D["y"]["h_data"].cpu().numpy().tobytes() == (-np.conj(vy)).tobytes()
    if N == 300:
        assert G[0].nnz == 2 * N * (N - 3)


def test_yeeder2d_bad_args():
    import fdfd_b200
    with pytest.raises(ValueError):
        fdfd_b200.yeeder2d([0, 3], [1, 1], [0, 0])
    with pytest.raises(ValueError):
        fdfd_b200.yeeder2d([3, 3], [1, 1], [2, 0])
