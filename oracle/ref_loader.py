"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

The reference is pure Python and every solver module imports matplotlib and/or tkinter at
module top (Scattering_Solver_2D.py:1, Band_Diagram_Solver.py:17-21, Mode_Solver_2D.py:1-7,
Periodic_Solver_3D.py:1-7); neither is installed here, so do-nothing stubs from
``oracle/stubs`` are put on ``sys.path`` first.  Nothing of the reference is copied into this
repository: the modules are loaded from where they lie, and only when /root/reference exists
(i.e. in the build container — the GPU box never has it; tests that need it skip there and
the committed fixtures under tests/golden/ stand in).

Only tests/, tests/golden/make_golden.py and oracle self-checks may import this module.
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("FDFD_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")

_FOLDER_OF = {
    "Scattering_Solver_2D": "Scattering",
    "Band_Diagram_Solver": "Band_Diagram_Solver",
    "Mode_Solver_2D": "Mode_Solver_2D",
    "Periodic_Solver_3D": "Periodic_Solver_3D",
    "refined_shift_invert_arnoldi": "Periodic_Solver_3D",
    "Periodic_Solver_2D": "Periodic_Solver_2D",
}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Scattering"))


def _ensure_stubs():
    try:  # a real matplotlib wins if one is ever installed
        import matplotlib  # noqa: F401
        import tkinter  # noqa: F401
    except Exception:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)


def load(module: str, folder: str | None = None):
    """Import reference module ``module`` (bare name, as the reference itself imports them)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _ensure_stubs()
    folder = folder or _FOLDER_OF.get(module)
    if folder is None:
        raise KeyError(module)
    path = os.path.join(REFERENCE_ROOT, folder)
    # yee_derivative exists in two folders with identical yeeder2d; keep them apart.
    for shared in ("yee_derivative",):
        sys.modules.pop(shared, None)
    sys.path.insert(0, path)
    try:
        sys.modules.pop(module, None)
        return importlib.import_module(module)
    finally:
        sys.path.remove(path)


def ref_yeeder2d(folder: str = "Scattering"):
    sys.modules.pop("yee_derivative", None)
    return load("yee_derivative", folder).yeeder2d
