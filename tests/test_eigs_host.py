"""Host-runnable checks of the eigen-solver algebra that is written with torch tensor ops (the same
code runs on the device in the product: batched Krylov-Schur of the Bloch sweep, folded
block-tridiagonal direct solver of the 3-D pencil).  These run the torch code on CPU tensors; the
parity tests proper are the GPU tests against the reference goldens."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla


def test_krylov_schur_batched_matches_dense_eig():
    import torch
    from fdfd_b200 import eigs
    rng = np.random.default_rng(0)
    P, n, k = 6, 90, 4
    mats = []
    for p in range(P):
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        mats.append(G @ G.conj().T / n + (0.01 + 0.02 * p) * np.eye(n))
    inv = torch.from_numpy(np.stack([np.linalg.inv(M) for M in mats]))
    state = {"inv": inv}

    def set_active(idx):
        state["inv"] = inv if idx is None else inv.index_select(0, idx)

    def op(W):
        return torch.bmm(state["inv"], W.unsqueeze(2)).squeeze(2)

    mu, X, res, restarts = eigs.krylov_schur_batched(op, set_active, P, n, k, tol=1e-12, device="cpu")
    assert restarts >= 1                      # the active set shrinks over several restart cycles
    for p in range(P):
        ref = np.sort(np.linalg.eigvalsh(mats[p]))[:k]
        got = np.sort((1 / mu[p]).real)
        assert np.max(np.abs(got - ref) / ref) <= 1e-10
        x = X[p].numpy()
        for i in range(k):
            assert np.linalg.norm(mats[p] @ x[i] - (1 / mu[p][i]) * x[i]) <= 1e-9


@pytest.mark.parametrize("Nz", [4, 5, 8])
def test_plane_block_solver_algebra(Nz):
    import torch
    import fdfd_b200
    from fdfd_b200 import eigs
    s = fdfd_b200.PeriodicModeSolver3D(6, 5, Nz, 3e-3, 2.5e-3, 2e-3, 30e9, 1, sigma_guess=-200j, tol=1e-8, ncv=10)
    s.add_block(4, 1, (1e-3, 2e-3), (0, 1e-3), (0, Nz), subpixels=2)
    s.add_pec((0, 3e-3), (2e-3, 2.5e-3), (0, Nz))
    # the class assembles its staggered operators on the GPU; on this CPU box the oracle's (bit-identical,
    # tests/test_oracle_vs_reference.py) stand in so that the host algebra of the block solver can be checked
    from oracle import periodic3d_oracle
    for name, mat in periodic3d_oracle.operators(6, 5, Nz, s.dx, s.dy, s.dz).items():
        setattr(s, name, mat)
    A, B = s._build_eigen_matrices()
    solver = eigs.PlaneBlockShiftedSolver(A.tocsr(), B.tocsr(), s.sigma_guess, s._unknown_planes(), s.Nz, "cpu", refine=False)
    assert solver.ns == (Nz + 1) // 2
    rng = np.random.default_rng(1)
    b = rng.standard_normal(A.shape[0]) + 1j * rng.standard_normal(A.shape[0])
    x = solver(torch.from_numpy(b)).numpy()
    S = (A - s.sigma_guess * B).tocsc()
    assert np.linalg.norm(S @ x - b) <= 1e-9 * np.linalg.norm(b)
    assert np.linalg.norm(x - spla.splu(S).solve(b)) <= 1e-8 * np.linalg.norm(x)
    with pytest.raises(ValueError):
        eigs.PlaneBlockShiftedSolver(A.tocsr(), B.tocsr(), s.sigma_guess, (s._unknown_planes() * 2) % Nz, Nz, "cpu", refine=False)


@pytest.mark.parametrize("pol", ["TE", "TM"])
@pytest.mark.parametrize("shape,bc", [((10, 15), (1, 1)), ((8, 9), (1, 1)), ((7, 11), (1, 1)), ((12, 7), (0, 0)),
                                      ((9, 10), (1, 0)), ((3, 4), (1, 1)), ((2, 5), (1, 1))])
def test_probed_dense_assembly_matches_explicit_matrix(pol, shape, bc):
    """Coloured probing of a five-point operator (BandDiagramSolver2D._dense_probed) against the explicit sparse
    matrix of the same operator; the operator is a stub that applies the oracle's matrix to CPU tensors."""
    import scipy.sparse as sp
    import torch
    import fdfd_b200
    from oracle import yee_oracle
    Nx, Ny = shape
    rng = np.random.default_rng(Nx * 100 + Ny)
    ca, cb, cd = (rng.uniform(0.5, 2, (Ny, Nx)) + 1j * rng.uniform(-1, 1, (Ny, Nx)) for _ in range(3))
    kinc = rng.standard_normal(2)
    DEX, DEY, DHX, DHY = yee_oracle.yeeder2d([Nx, Ny], [1.0, 1.0], list(bc), kinc)
    a, b, d = sp.diags(ca.ravel()), sp.diags(cb.ravel()), sp.diags(cd.ravel())
    M = ((DHX @ a @ DEX) + (DHY @ b @ DEY) + d if pol == "TE" else (DEX @ a @ DHX) + (DEY @ b @ DHY) + d).tocsr()

    class StubOperator:
        n, tdtype, device = Nx * Ny, torch.complex128, torch.device("cpu")
        applies = 0

        def apply(self, x, out=None):
            StubOperator.applies += 1
            y = torch.from_numpy(M @ x.numpy())
            if out is None:
                return y
            out.copy_(y)
            return out
    op = StubOperator()
    op.Nx, op.Ny, op.bc = Nx, Ny, bc
    solver = fdfd_b200.BandDiagramSolver2D.__new__(fdfd_b200.BandDiagramSolver2D)
    D = solver._dense_probed(op).numpy()
    assert np.array_equal(D, M.toarray())
    px = solver._probe_period(Nx, bc[0] == 1)
    py = solver._probe_period(Ny, bc[1] == 1)
    if px is not None and py is not None:
        assert StubOperator.applies == px * py <= Nx * Ny
    else:
        assert StubOperator.applies == Nx * Ny          # too short to colour: one apply per column


def test_band_projected_problems_do_not_depend_on_the_thread_cut():
    """BandDiagramSolver2D._projected_problems (thread pool over chunks of path points) equals the single stacked
    call bit for bit, and its Ritz values / restart blocks satisfy the Krylov-Schur relations."""
    from fdfd_b200.band_diagram import BandDiagramSolver2D as B
    rng = np.random.default_rng(3)
    P, ncv, k, keep = 100, 20, 5, 12
    G = rng.standard_normal((P, ncv, ncv)) + 1j * rng.standard_normal((P, ncv, ncv))
    Hm = G @ np.conj(G.transpose(0, 2, 1)) / ncv
    b = rng.standard_normal((P, ncv)) + 1j * rng.standard_normal((P, ncv))
    one = B._projected_chunk(Hm, b, k, keep)
    many = B._projected_problems(Hm, b, k, keep, min_chunk=7)
    assert all(np.array_equal(x, y) for x, y in zip(one, many))
    mu, res, Sk, Qk, Tn = many
    for p in (0, 57, 99):
        assert np.allclose(np.sort(np.abs(mu[p]))[::-1], np.abs(mu[p]))                 # dominant first
        for i in range(k):
            assert np.linalg.norm(Hm[p] @ Sk[p, i] - mu[p, i] * Sk[p, i]) <= 1e-10 * abs(mu[p, i])
        Q = Qk[p].T
        assert np.allclose(Q.conj().T @ Q, np.eye(keep), atol=1e-12)
        assert np.allclose(Tn[p, :keep, :keep], Q.conj().T @ Hm[p] @ Q, atol=1e-12)
        assert np.allclose(Tn[p, keep, :keep], b[p] @ Q, atol=1e-12)


def test_ruiz_equilibrate_scales_rows_and_columns_to_one():
    import scipy.sparse as sp
    from fdfd_b200 import eigs
    rng = np.random.default_rng(5)
    n = 300
    M = sp.random(n, n, density=0.02, random_state=5, format="csr") + sp.identity(n)
    M = sp.diags(10.0 ** rng.uniform(-6, 6, n)) @ M @ sp.diags(10.0 ** rng.uniform(-6, 6, n))     # 24 decades of scale
    dr, dc = eigs.ruiz_equilibrate(M, iterations=12)
    S = abs(sp.diags(dr) @ M @ sp.diags(dc)).tocsr()
    rmax, cmax = np.asarray(S.max(axis=1).todense()).ravel(), np.asarray(S.max(axis=0).todense()).ravel()
    assert np.all(np.abs(rmax - 1) <= 0.05) and np.all(np.abs(cmax - 1) <= 0.05)
    assert np.all(dr > 0) and np.all(dc > 0)
