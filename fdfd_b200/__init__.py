"""fdfd_b200 — B200-native (sm_100a) solve engine behind the FDFD reference API.

Public surface (names as in SolverNotConverging/FDFD):
    yeeder2d(NS, RES, BC, kinc=None)          GPU assembly of the Yee derivative operators
    FDFD2DScatteringSolver                    2-D TEz/TMz scattering, matrix-free GPU Krylov solve
    BandDiagramSolver2D                       photonic band structure, GPU Krylov-Schur shift-invert
    ModeSolver2D, PeriodicModeSolver2D,
    PeriodicModeSolver3D                      waveguide / periodic-cell eigenmodes, GPU shift-invert Arnoldi
    HelmholtzOperator2D                       the matrix-free operator + solver handle
"""
from ._lib import FdfdError, NoConvergence, LIB_PATH  # noqa: F401
from .yee_derivative import yeeder2d, yeeder2d_device  # noqa: F401
from .operators import HelmholtzOperator2D, SlabComm, slab_rows  # noqa: F401
from .scattering import FDFD2DScatteringSolver  # noqa: F401
from .band_diagram import BandDiagramSolver2D, BandStructureResult  # noqa: F401
from .mode_solver_2d import ModeSolver2D  # noqa: F401
from .periodic_solver_3d import PeriodicModeSolver3D  # noqa: F401
from .periodic_solver_2d import PeriodicModeSolver2D  # noqa: F401
from .eigs import refined_shift_invert_arnoldi  # noqa: F401
from . import sweeps, stagger  # noqa: F401


def trim_memory():
    """Release the library's cached device work buffers (csrc/devpool.cu); returns the bytes released."""
    from . import _lib
    return int(_lib.lib().fdfd_trim_memory())

__all__ = ["yeeder2d", "yeeder2d_device", "HelmholtzOperator2D", "SlabComm", "slab_rows",
           "FDFD2DScatteringSolver", "BandDiagramSolver2D", "BandStructureResult", "ModeSolver2D", "PeriodicModeSolver3D", "PeriodicModeSolver2D", "refined_shift_invert_arnoldi", "sweeps", "stagger", "FdfdError", "NoConvergence", "trim_memory"]
