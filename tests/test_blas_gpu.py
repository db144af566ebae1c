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
