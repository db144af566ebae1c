"""K1b — staggered-grid difference / average operators assembled on the GPU (csrc/yee_assemble.cu,
C-ABI ``fdfd_stagger_csr_fill*``), returned as the scipy objects the reference builds with Python loops:
Mode_Solver_2D/Mode_Solver_2D.py:482-544 and Periodic_Solver_3D/Periodic_Solver_3D.py:95-194.

Every row has two entries (self and forward neighbour); the kernel writes ``indptr = 2*row`` and the sorted
column pair in one pass.  The two values are evaluated HERE with the reference's numpy/Python expressions
(``val / scale``), so data as well as indices are bit-identical to scipy's ``coo -> csr`` result."""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib


def _shape3(s):
    s = tuple(int(v) for v in s)
    return s + (1,) * (3 - len(s))


def stagger_csr(in_shape, out_shape, axis, v_self, v_fwd, with_h=False):
    """CSR operator (rows = output grid, columns = input grid), float64; with_h=True also returns the CSC
    matrix ``-D^H`` (same index arrays, negated data), as the reference forms its H-field operators."""
    _lib.require_cuda()
    L = _lib.lib()
    ins, outs = _shape3(in_shape), _shape3(out_shape)
    rows, cols = int(np.prod(outs)), int(np.prod(ins))
    indptr = np.empty(rows + 1, dtype=np.int32)
    indices = np.empty(2 * rows, dtype=np.int32)
    data = np.empty(2 * rows, dtype=np.float64)
    hdata = np.empty(2 * rows, dtype=np.float64) if with_h else None
    _lib.check(L.fdfd_stagger_csr_fill_host((C.c_int * 3)(*ins), (C.c_int * 3)(*outs), int(axis), float(v_self), float(v_fwd),
                                            indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                                            None if hdata is None else hdata.ctypes.data))
    D = sp.csr_matrix((data, indices, indptr), shape=(rows, cols))
    if not with_h:
        return D
    H = sp.csc_matrix((hdata, indices.copy(), indptr.copy()), shape=(cols, rows))
    return D, H


def difference(in_shape, out_shape, scale, axis, with_h=False):
    """Forward difference along `axis` scaled by 1/scale (reference: `val / scale` for val = +1, -1)."""
    return stagger_csr(in_shape, out_shape, axis, -1.0 / scale, 1.0 / scale, with_h=with_h)


def periodic_z(shape, forward_weight, self_weight):
    """`forward_weight * f[(k+1) % nz] + self_weight * f[k]` on one grid (Periodic_Solver_3D.py:161-194)."""
    return stagger_csr(shape, shape, 2, self_weight, forward_weight)
