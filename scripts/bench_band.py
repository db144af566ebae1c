"""BASELINE config 3 (example_square_lattice.py) end to end on the GPU, and the oracle (scipy) timing."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 200
s = fdfd_b200.BandDiagramSolver2D(a=1.0, Nx=40, background_er=10.2)
s.add_circular_inclusion(radius=0.4, er=1.0)
pts, labels = s.default_rectangular_lattice_path()
bp, ticks = s.generate_bloch_path(pts, total_points=npts)
t = time.time()
r = s.compute_band_structure(bp, num_bands=5)
first = time.time() - t            # includes CUDA context / cuSOLVER / cuBLAS initialisation of the process
t = time.time()
r = s.compute_band_structure(bp, num_bands=5)
dt = time.time() - t
out = dict(config="square lattice 40x40, %d Bloch points x 2 polarisations, 5 bands" % npts, gpu_seconds=round(dt, 3), gpu_seconds_first_call=round(first, 3), **s.solver_info)
if "--cpu" in sys.argv:
    from oracle import band_oracle as bo
    p = bo.BandProblem(1.0, 40, background_er=10.2)
    p.add_circle(0.4, er=1.0)
    t = time.time()
    o = bo.band_structure(p, bp, 5)
    out["cpu_scipy_seconds"] = round(time.time() - t, 3)
    out["max_rel_eig_diff"] = float(max(np.max(np.abs(r.eigenvalues[q] - o[q][0]) / np.maximum(np.abs(o[q][0]), 1e-2)) for q in ("TE", "TM")))
print(json.dumps(out))
