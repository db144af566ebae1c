"""K4 multi-vector kernels (Gram-Schmidt building blocks) against torch: even / odd lengths (pair path with
tail, scalar fallback for odd strides), complex128 and complex64, more vectors than one pass holds;
lengths >= 2^20 go through the bulk-copy (TMA) staged kernels, with and without a ragged tail."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-12), ("complex64", 2e-5)])
@pytest.mark.parametrize("n", [1, 2, 1000, 1001, 65537, 262144, 1048576 + 1000, 2097153, 3 * 1048576])
@pytest.mark.parametrize("k", [1, 3, 8, 11])
def test_multi_dot_axpy(dtype, tol, n, k):
    import torch
    from fdfd_b200.eigs import DeviceBlas
    blas = DeviceBlas("cuda")
    tdt = getattr(torch, dtype)
    g = torch.Generator(device="cuda").manual_seed(n * 31 + k)
    V = torch.randn((k, n), dtype=tdt, device="cuda", generator=g)
    w = torch.randn(n, dtype=tdt, device="cuda", generator=g)
    h = blas.multi_dot(V, k, w)
    ref = (V.conj().to(torch.complex128) @ w.to(torch.complex128)).cpu().numpy()
    assert np.allclose(h, ref, rtol=tol, atol=tol * np.sqrt(n))
    w2 = w.clone()
    nrm2 = blas.multi_axpy(V, k, w2, h, -1.0, want_norm=True)
    ref_w = w.to(torch.complex128) - torch.from_numpy(h).cuda() @ V.to(torch.complex128)
    assert float(torch.linalg.vector_norm(w2.to(torch.complex128) - ref_w)) <= 10 * tol * max(1.0, float(torch.linalg.vector_norm(w)) * np.sqrt(k))
    assert abs(nrm2 - float(torch.linalg.vector_norm(w2.to(torch.complex128)) ** 2)) <= 1e-4 * max(nrm2, 1e-30) if dtype == "complex64" \
        else abs(nrm2 - float(torch.linalg.vector_norm(w2) ** 2)) <= 1e-10 * max(nrm2, 1e-30)


@pytest.mark.parametrize("n,m", [(50, 7), (300, 40), (10007, 24), (119040, 40), (5000, 64)])
def test_tsqr_r_matches_lapack_on_nearly_singular_matrices(n, m):
    """K7 (csrc/tsqr.cu): R of (AV - theta BV)^T.  The matrices of the refined extraction are singular to working
    precision on purpose, so the check is on the singular values (down to the smallest) and on the right singular
    vector of the smallest one, against numpy's SVD of the explicit n x m matrix."""
    import torch
    from fdfd_b200.eigs import DeviceBlas
    rng = np.random.default_rng(n + m)
    Q1 = np.linalg.qr(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m)))[0]
    Q2 = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))[0]
    sv = np.logspace(0, -8, m)
    sv[-1] = 1e-11                                          # singular values 1 ... 1e-8 and one isolated at 1e-11
    W = (Q1 * sv) @ Q2.conj().T
    theta = 0.3 - 0.7j
    BV = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    AV = W.T + theta * BV                                    # so that AV - theta BV = W^T
    W = (AV - theta * BV).T                                  # what the kernel sees (rounded)
    blas = DeviceBlas()
    R = blas.tsqr_r(torch.from_numpy(AV).cuda(), torch.from_numpy(BV).cuda(), theta)
    assert R.shape == (m, m) and np.allclose(np.tril(R, -1), 0)
    s_ref = np.linalg.svd(W, compute_uv=False)
    s_got = np.linalg.svd(R, compute_uv=False)
    assert np.max(np.abs(s_got - s_ref)) <= 1e-13 * s_ref[0]   # 1 % of the smallest singular value
    # Gram matrices agree: R^H R = W^H W
    G = W.conj().T @ W
    assert np.linalg.norm(R.conj().T @ R - G) <= 1e-12 * np.linalg.norm(G)
    # R alone (no BV)
    R1 = blas.tsqr_r(torch.from_numpy(np.ascontiguousarray(W.T)).cuda())
    assert np.linalg.norm(R1.conj().T @ R1 - G) <= 1e-13 * np.linalg.norm(G)
    y_ref = np.linalg.svd(W, full_matrices=False)[2].conj().T[:, -1]
    y_got = np.linalg.svd(R1)[2].conj().T[:, -1]
    assert abs(abs(np.vdot(y_ref, y_got)) - 1) <= 1e-6        # angle ~ 1e-16 sqrt(n m) / 1e-8, enters squared
    blas.close()
