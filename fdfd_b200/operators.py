"""Host handle of the matrix-free Helmholtz / curl-curl operator K2 and of the Krylov stack.

Mirrors what the reference materialises as scipy sparse matrices:
  ``A_TE = DHX mu_yy^-1 DEX + DHY mu_xx^-1 DEY + eps_zz``  (Scattering_Solver_2D.py:176-181)
  ``A_TM = DEX eps_yy^-1 DHX + DEY eps_xx^-1 DHY + mu_zz`` (Scattering_Solver_2D.py:183-188)
  ``A_tm = -DHX mu_yy^-1 DEX - DHY mu_xx^-1 DEY``          (Band_Diagram_Solver.py:312)
  ``A_te = -DEX eps_yy^-1 DHX - DEY eps_xx^-1 DHY``        (Band_Diagram_Solver.py:319)
but never forms them: ``apply`` runs the sm_100a stencil kernel through the C-ABI.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib

_POL = {"TE": _lib.POL_TE, "TM": _lib.POL_TM}


def _torch():
    import torch
    return torch


class HelmholtzOperator2D:
    """y = sx * D1 diag(ca) D2 x + sy * D3 diag(cb) D4 x + cd .* x   on an Nx x Ny Yee grid.

    pol='TE': D = (DHX, DEX, DHY, DEY); pol='TM': D = (DEX, DHX, DEY, DHY)  (unit spacing).
    ca, cb, cd: (Ny, Nx) complex arrays (numpy or torch); cd may be None.
    bc: (xbc, ybc) with 0 Dirichlet / 1 Bloch; phase = exp(-1j*k*N*d) per axis.
    Sharding: ``rows=(j0, j1)`` selects this rank's y-slab, ``comm`` a :class:`SlabComm`.
    """

    def __init__(self, pol, Nx, Ny, ca, cb, cd, sx, sy, bc=(0, 0), phase=(1.0 + 0j, 1.0 + 0j),
                 dtype="complex128", device=None, rows=None, comm=None):
        torch = _torch()
        _lib.require_cuda()
        self._L = _lib.lib()
        self.pol = pol.upper()
        if self.pol not in _POL:
            raise ValueError("pol must be 'TE' or 'TM'")
        self.Nx, self.Ny = int(Nx), int(Ny)
        self.device = torch.device(device or "cuda")
        self.tdtype = {"complex128": torch.complex128, "complex64": torch.complex64}[str(dtype).replace("torch.", "")]
        self.code = _lib.C128 if self.tdtype == torch.complex128 else _lib.C64
        self.j0, self.j1 = (0, self.Ny) if rows is None else (int(rows[0]), int(rows[1]))
        self.ny_local = self.j1 - self.j0
        self.comm = comm
        shape = (self.ny_local, self.Nx)
        self.ca = self._dev(ca, shape)
        self.cb = self._dev(cb, shape)
        self.cd = None if cd is None else self._dev(cd, shape)
        self.sx, self.sy = float(sx), float(sy)
        self.bc = (int(bc[0]), int(bc[1]))
        ph = (C.c_double * 4)(complex(phase[0]).real, complex(phase[0]).imag,
                              complex(phase[1]).real, complex(phase[1]).imag)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_helm2d_create(
                C.byref(self._h), self.code, _POL[self.pol], self.Nx, self.Ny, self.j0, self.ny_local,
                self.bc[0], self.bc[1], ph, self.sx, self.sy, self.ca.data_ptr(), self.cb.data_ptr(),
                None if self.cd is None else self.cd.data_ptr(),
                None if comm is None else comm.handle, self._stream()))
        if comm is not None:
            comm._register(self)

    # -- helpers ------------------------------------------------------------------------------
    def _dev(self, a, shape):
        torch = _torch()
        if not isinstance(a, torch.Tensor):
            a = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.complex128).reshape(shape)))
        a = a.to(device=self.device, dtype=self.tdtype).reshape(shape).contiguous()
        return a

    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    @property
    def n(self):
        return self.Nx * self.ny_local

    @property
    def bytes_per_apply(self):
        return int(self._L.fdfd_helm2d_bytes_per_apply(self._h))

    def invalidate(self):
        """Drop the cached preconditioner (after changing the coefficient tensors in place)."""
        _lib.check(self._L.fdfd_helm2d_invalidate(self._handle()))

    def set_variant(self, v):
        _lib.check(self._L.fdfd_helm2d_set_variant(self._h, int(v)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.fdfd_helm2d_destroy(self._h)
            self._h = C.c_void_p()

    def _release_comm_state(self):
        """Called by `SlabComm.close()`: a sharded operator owns ghost-row buffers in the communicator's
        peer-mapped heap (and, on the NCCL fallback path, a captured graph that pins the communicator), so
        it is closed together with it; using it afterwards raises instead of touching freed memory."""
        self.close()

    def _handle(self):
        if getattr(self, "_h", None) is None or not self._h.value:
            raise _lib.FdfdError("operator is closed (its communicator was closed or close() was called)")
        return self._h

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- operator application -------------------------------------------------------------------
    def apply(self, x, out=None):
        """y = A x for a device tensor x (any shape with n elements)."""
        torch = _torch()
        if not isinstance(x, torch.Tensor):
            raise TypeError("apply() takes a CUDA torch tensor; use apply_host() for numpy arrays")
        if x.numel() != self.n or x.dtype != self.tdtype or not x.is_cuda:
            raise ValueError("x must be a CUDA tensor with Nx*Ny_local elements of the operator dtype")
        x = x.contiguous()
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_helm2d_apply(self._handle(), x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def apply_host(self, x, out=None):
        """y = A x with numpy in / numpy out (H2D + kernel + D2H inside the C-ABI call).
        `out` may be a preallocated (e.g. pinned) array of the same dtype and size."""
        ndt = np.complex128 if self.code == _lib.C128 else np.complex64
        x = np.ascontiguousarray(x, dtype=ndt)
        if x.size != self.n:
            raise ValueError("x has the wrong number of elements")
        y = np.empty_like(x) if out is None else out
        if y.dtype != ndt or y.size != self.n or not y.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous array of the operator dtype and size")
        with _torch().cuda.device(self.device):
            _lib.check(self._L.fdfd_helm2d_apply_host(self._handle(), x.ctypes.data, y.ctypes.data))
        return y

    __matmul__ = lambda self, x: self.apply(x) if not isinstance(x, np.ndarray) else self.apply_host(x)

    # -- linear solve -------------------------------------------------------------------------------
    def solve(self, b, x0=None, *, method="bicgstab", precond="jacobi", rtol=1e-8, maxit=100000,
              raise_on_noconv=True, **opts):
        """Solve A x = b on the device.  Returns (x, info dict)."""
        torch = _torch()
        o = _lib.default_opts(method={"bicgstab": _lib.KRYLOV_BICGSTAB, "fgmres": _lib.KRYLOV_FGMRES}[method],
                              precond={"none": _lib.PRECOND_NONE, None: _lib.PRECOND_NONE,
                                       "jacobi": _lib.PRECOND_JACOBI, "mg": _lib.PRECOND_MG}[precond],
                              rtol=float(rtol), maxit=int(maxit), **opts)
        info = (C.c_double * 8)()
        host = isinstance(b, np.ndarray)
        ndt = np.complex128 if self.code == _lib.C128 else np.complex64
        if host:
            bh = np.ascontiguousarray(b, dtype=ndt).ravel()
            xh = np.zeros_like(bh) if x0 is None else np.ascontiguousarray(x0, dtype=ndt).ravel().copy()
            st = self._L.fdfd_krylov_solve_host(self._handle(), C.byref(o), bh.ctypes.data, xh.ctypes.data, info)
            x = xh
        else:
            bd = b.to(device=self.device, dtype=self.tdtype).contiguous().reshape(-1)
            x = torch.zeros_like(bd) if x0 is None else x0.to(device=self.device, dtype=self.tdtype).reshape(-1).clone()
            with torch.cuda.device(self.device):
                st = self._L.fdfd_krylov_solve(self._handle(), C.byref(o), bd.data_ptr(), x.data_ptr(), info, self._stream())
        d = dict(iterations=int(info[0]), relres=float(info[1]), applies=int(info[2]),
                 precond_applies=int(info[3]), seconds=float(info[4]), breakdowns=int(info[5]),
                 method=method, precond=precond, converged=(st == _lib.FDFD_OK))
        if st == _lib.FDFD_ERR_NOCONV and not raise_on_noconv:
            return x, d
        _lib.check(st, d)
        return x, d


    def precond_apply(self, v, precond="mg", **opts):
        """z = M^-1 v (one preconditioner application) — validation hook for the tests."""
        torch = _torch()
        o = _lib.default_opts(precond={"jacobi": _lib.PRECOND_JACOBI, "mg": _lib.PRECOND_MG}[precond], **opts)
        v = v.to(device=self.device, dtype=self.tdtype).contiguous().reshape(-1)
        z = torch.empty_like(v)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_precond_apply(self._handle(), C.byref(o), v.data_ptr(), z.data_ptr(), self._stream()))
        return z


class CsrProductOperator:
    """y = P (Q x) + shift * x with P (n x m) and Q (m x n) scipy CSR matrices uploaded once — the
    mode-solver operator Omega = P_red @ Q_red (Mode_Solver_2D.py:771-774) kept factored on the GPU.
    ``solve`` applies (P Q + shift I)^-1 with the K3 Krylov solvers (Jacobi: 1/(diag(PQ)+shift))."""

    def __init__(self, P, Q, shift=0.0, device=None, diag=None):
        torch = _torch()
        _lib.require_cuda()
        self._L = _lib.lib()
        self.device = torch.device(device or "cuda")
        P = P.tocsr()
        Q = Q.tocsr()
        if P.shape[1] != Q.shape[0] or P.shape[0] != Q.shape[1]:
            raise ValueError("P and Q shapes do not form a square product")
        self.n, self.m = P.shape
        self.tdtype = torch.complex128
        self.shift = complex(shift)

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        self._keep = [up(P.indptr, np.int32), up(P.indices, np.int32), up(P.data, np.complex128),
                      up(Q.indptr, np.int32), up(Q.indices, np.int32), up(Q.data, np.complex128)]
        if diag is None:
            diag = np.asarray(P.multiply(Q.T).sum(axis=1)).ravel()
        self.diag = np.asarray(diag, dtype=np.complex128)
        d = self.diag + self.shift
        d = np.where(np.abs(d) > 0, d, 1.0)          # empty rows (fully constrained neighbours): identity
        self.dinv = up(1.0 / d, np.complex128)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_csr_product_create(C.byref(self._h), _lib.C128, self.n, self.m,
                                                       *[t.data_ptr() for t in self._keep],
                                                       self.shift.real, self.shift.imag))

    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.fdfd_linop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, x, out=None):
        torch = _torch()
        x = x.to(device=self.device, dtype=self.tdtype).contiguous().reshape(-1)
        if x.numel() != self.n:
            raise ValueError("x has the wrong number of elements")
        if out is None:
            out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_linop_apply(self._h, x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def solve(self, b, x0=None, *, method="bicgstab", rtol=1e-12, maxit=20000, jacobi=True, raise_on_noconv=True,
              **opts):
        torch = _torch()
        o = _lib.default_opts(method={"bicgstab": _lib.KRYLOV_BICGSTAB, "fgmres": _lib.KRYLOV_FGMRES}[method],
                              rtol=float(rtol), maxit=int(maxit), **opts)
        b = b.to(device=self.device, dtype=self.tdtype).contiguous().reshape(-1)
        x = torch.zeros_like(b) if x0 is None else x0.to(device=self.device, dtype=self.tdtype).reshape(-1).clone()
        info = (C.c_double * 8)()
        with torch.cuda.device(self.device):
            st = self._L.fdfd_linop_solve(self._h, C.byref(o), self.dinv.data_ptr() if jacobi else None,
                                          b.data_ptr(), x.data_ptr(), info, self._stream())
        d = dict(iterations=int(info[0]), relres=float(info[1]), applies=int(info[2]), seconds=float(info[4]),
                 converged=(st == _lib.FDFD_OK))
        if st == _lib.FDFD_ERR_NOCONV and not raise_on_noconv:
            return x, d
        _lib.check(st, d)
        return x, d


class CsrOperator:
    """y = A x (+ shift x) for one scipy CSR matrix uploaded to the device (rectangular allowed when shift = 0)."""

    def __init__(self, A, shift=0.0, device=None):
        torch = _torch()
        _lib.require_cuda()
        self._L = _lib.lib()
        self.device = torch.device(device or "cuda")
        A = A.tocsr()
        self.shape = tuple(int(v) for v in A.shape)
        if A.shape[0] != A.shape[1] and complex(shift) != 0:
            raise ValueError("a shift needs a square matrix")
        self.n = A.shape[1]            # input length
        self.nrows = A.shape[0]
        self.tdtype = torch.complex128

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        self._keep = [up(A.indptr, np.int32), up(A.indices, np.int32), up(A.data, np.complex128)]
        self._h = C.c_void_p()
        sh = complex(shift)
        with torch.cuda.device(self.device):
            if self.shape[0] == self.shape[1]:
                _lib.check(self._L.fdfd_csr_create(C.byref(self._h), _lib.C128, self.n, *[t.data_ptr() for t in self._keep],
                                                   sh.real, sh.imag))
            else:
                _lib.check(self._L.fdfd_csr_create_rect(C.byref(self._h), _lib.C128, self.nrows, self.n,
                                                        *[t.data_ptr() for t in self._keep]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.fdfd_linop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, x, out=None):
        torch = _torch()
        x = x.to(device=self.device, dtype=self.tdtype).contiguous().reshape(-1)
        if x.numel() != self.n:
            raise ValueError("x has the wrong number of elements")
        if out is None:
            out = torch.empty(self.nrows, dtype=self.tdtype, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.fdfd_linop_apply(self._h, x.data_ptr(), out.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def apply_rows(self, V):
        """A applied to every row of V (m, n) -> (m, nrows): the SpMM of the refined extraction."""
        torch = _torch()
        out = torch.empty((V.shape[0], self.nrows), dtype=self.tdtype, device=self.device)
        for i in range(V.shape[0]):
            self.apply(V[i], out=out[i])
        return out


class SlabComm:
    """NCCL communicator for y-slab sharding; the 128-byte unique id travels through
    ``torch.distributed`` (any backend) from rank 0."""

    def __init__(self, rank=None, world_size=None):
        import torch
        import torch.distributed as dist

        L = _lib.lib()
        self.rank = dist.get_rank() if rank is None else rank
        self.world_size = dist.get_world_size() if world_size is None else world_size
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(L.fdfd_comm_unique_id(buf))
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        ident = bytes(t.cpu().numpy().tobytes())
        self.handle = C.c_void_p()
        _lib.check(L.fdfd_comm_create(C.byref(self.handle), self.world_size, self.rank, ident))
        self._L = L
        self._dependants = weakref.WeakSet()

    def _register(self, op):
        self._dependants.add(op)

    @property
    def peer_mapped(self):
        """True when ghost rows / small collectives travel over NVLink peer memory (CUDA IPC heap), False on
        the NCCL fallback path."""
        return bool(self.handle and self.handle.value and self._L.fdfd_comm_peer_mapped(self.handle))

    def close(self):
        """Destroy the communicator.  Operators created with it drop their cached preconditioners
        first: a captured multigrid cycle references the communicator (NCCL keeps a persistent
        reference per captured graph) and ncclCommDestroy would otherwise wait for ever."""
        if self.handle and self.handle.value:
            import torch
            for op in list(self._dependants):
                op._release_comm_state()
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self._L.fdfd_comm_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def slab_rows(Ny, world_size, rank):
    """Contiguous y-slab [j0, j1) of rank `rank`; remainders go to the low ranks."""
    base, rem = divmod(int(Ny), int(world_size))
    j0 = rank * base + min(rank, rem)
    return j0, j0 + base + (1 if rank < rem else 0)
