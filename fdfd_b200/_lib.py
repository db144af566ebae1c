"""ctypes binding of libfdfd_b200.so (the C-ABI declared in include/fdfd_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
visible, every compute entry point raises.  The oracle under ``oracle/`` is test
infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDFD_B200_LIB") or os.path.join(_HERE, "libfdfd_b200.so")   # override: debug builds only

FDFD_OK = 0
FDFD_ERR_ARG, FDFD_ERR_CUDA, FDFD_ERR_BREAKDOWN, FDFD_ERR_NOCONV, FDFD_ERR_NCCL, FDFD_ERR_STATE = -1, -2, -3, -4, -5, -6
C128, C64 = 0, 1
POL_TE, POL_TM = 0, 1
BC_DIRICHLET, BC_BLOCH = 0, 1
KRYLOV_BICGSTAB, KRYLOV_FGMRES = 0, 1
PRECOND_NONE, PRECOND_JACOBI, PRECOND_MG = 0, 1, 2


class FdfdError(RuntimeError):
    pass


class NoConvergence(FdfdError):
    """Raised when the Krylov solver hits maxit (mirrors scipy's ArpackNoConvergence idea)."""

    def __init__(self, msg, info=None):
        super().__init__(msg)
        self.info = info


class KrylovOpts(C.Structure):
    _fields_ = [
        ("method", C.c_int), ("precond", C.c_int), ("rtol", C.c_double), ("maxit", C.c_int),
        ("restart", C.c_int), ("mg_levels", C.c_int), ("mg_pre", C.c_int), ("mg_post", C.c_int),
        ("mg_shift_re", C.c_double), ("mg_shift_im", C.c_double), ("mg_omega", C.c_double),
        ("mg_cycle", C.c_int), ("mg_coarse_its", C.c_int), ("verbose", C.c_int),
        ("line_x_lo", C.c_int), ("line_x_hi", C.c_int), ("line_y_lo", C.c_int), ("line_y_hi", C.c_int),
        ("mg_wfrom", C.c_int), ("mg_min_n", C.c_int), ("reorth", C.c_int),
        ("mg_coarse_mode", C.c_int), ("mg_coarse_replicate", C.c_int), ("mg_single", C.c_int), ("mg_graph", C.c_int), ("mg_kh_max", C.c_double), ("mg_fuse_len", C.c_int), ("basis_single", C.c_int),
    ]


_lib = None

_PROTOTYPES = {
    "fdfd_last_error": (C.c_char_p, []),
    "fdfd_version": (C.c_int, []),
    "fdfd_krylov_opts_size": (C.c_int, []),
    "fdfd_device_count": (C.c_int, []),
    "fdfd_trim_memory": (C.c_int64, []),
    "fdfd_yee2d_csr_nnz": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "fdfd_yee2d_csr_fill": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_yee2d_csr_fill_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_scatter_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "fdfd_stagger_csr_fill": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_stagger_csr_fill_host": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_helm2d_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_helm2d_destroy": (None, [C.c_void_p]),
    "fdfd_helm2d_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_helm2d_apply_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_helm2d_bytes_per_apply": (C.c_int64, [C.c_void_p]),
    "fdfd_helm2d_invalidate": (C.c_int, [C.c_void_p]),
    "fdfd_helm2d_set_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "fdfd_krylov_default_opts": (None, [C.POINTER(KrylovOpts)]),
    "fdfd_krylov_solve": (C.c_int, [C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
    "fdfd_krylov_solve_host": (C.c_int, [C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "fdfd_blas_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "fdfd_blas_destroy": (None, [C.c_void_p]),
    "fdfd_blas_multi_dot": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "fdfd_blas_multi_axpy": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "fdfd_tsqr_r": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p,
                              C.c_void_p]),
    "fdfd_band_max_nx": (C.c_int, []),
    "fdfd_band_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p]),
    "fdfd_band_destroy": (None, [C.c_void_p]),
    "fdfd_band_set_start": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_band_set_active": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fdfd_band_cycle": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fdfd_band_restart": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fdfd_band_ritz": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                 C.c_void_p]),
    "fdfd_band_vectors": (C.c_void_p, [C.c_void_p]),
    "fdfd_band_apply_shifted": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_csr_product_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "fdfd_csr_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_double]),
    "fdfd_csr_create_rect": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_linop_destroy": (None, [C.c_void_p]),
    "fdfd_linop_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_linop_solve": (C.c_int, [C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_double), C.c_void_p]),
    "fdfd_precond_apply": (C.c_int, [C.c_void_p, C.POINTER(KrylovOpts), C.c_void_p, C.c_void_p, C.c_void_p]),
    "fdfd_comm_unique_id": (C.c_int, [C.c_void_p]),
    "fdfd_comm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]),
    "fdfd_comm_peer_mapped": (C.c_int, [C.c_void_p]),
    "fdfd_comm_destroy": (None, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FdfdError(
                f"{LIB_PATH} not found — build it with `python -m fdfd_b200.build` "
                "(fdfd_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.fdfd_krylov_opts_size() != C.sizeof(KrylovOpts):
            raise FdfdError(f"{LIB_PATH} was built from another include/fdfd_b200.h (fdfd_krylov_opts is "
                            f"{handle.fdfd_krylov_opts_size()} bytes there, {C.sizeof(KrylovOpts)} here): rebuild it")
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().fdfd_last_error().decode("utf-8", "replace")


def check(status: int, info=None):
    if status == FDFD_OK:
        return
    msg = last_error()
    if status == FDFD_ERR_ARG:
        raise ValueError(msg)
    if status == FDFD_ERR_NOCONV:
        raise NoConvergence(msg, info)
    raise FdfdError(f"fdfd_b200 error {status}: {msg}")


def require_cuda():
    if lib().fdfd_device_count() <= 0:
        raise FdfdError("no CUDA device visible: fdfd_b200 computes on the GPU only (no CPU fallback)")


def default_opts(**kw) -> KrylovOpts:
    o = KrylovOpts()
    lib().fdfd_krylov_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solver option {k!r}")
        setattr(o, k, v)
    return o
