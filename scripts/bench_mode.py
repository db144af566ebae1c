"""BASELINE config 2 (example_ridge_dielectric_waveguide.py: 240 x 160 cells, 5 modes) end to end on the GPU."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402


def ridge():
    s = fdfd_b200.ModeSolver2D(50e9, 24e-3, 16e-3, 240, 160, 5)
    s.add_rectangle((3.0, 4.0, 5.0), 1, (0, 240), (60, 80))
    s.add_rectangle(6.0, 1, (100, 140), (80, 100))
    s.add_UPML(pml_width=30, n=3, sigma_max=1, direction="x")
    return s


reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
times = []
for _ in range(reps):                       # the first call includes CUDA context / library initialisation
    t = time.time()
    s = ridge()
    s.solve()
    times.append(round(time.time() - t, 3))
info = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in s.solver_info.items()}
print(json.dumps(dict(config="ridge waveguide 240x160, 5 modes", seconds_per_call=times, neff=[complex(z).real for z in s.neff],
                      **{k: (float(v) if hasattr(v, "real") and not isinstance(v, (list, int)) else v) for k, v in info.items()})))
