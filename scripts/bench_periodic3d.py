"""Time-to-solution of the 3-D periodic mode solver on BASELINE config 4 (and the sweep point)."""
import importlib.util
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import fdfd_b200  # noqa: E402

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "make_golden.py"))
m = importlib.util.module_from_spec(spec)
spec.loader.exec_module(m)

for name, build in (("sweep21", m.periodic3d_cases()["sweep21"]), ("cfg4", m.periodic3d_cfg4), ("cfg4", m.periodic3d_cfg4)):
    t0 = time.time()
    s, kw = build(fdfd_b200.PeriodicModeSolver3D)
    t1 = time.time()
    s.solve(**kw)
    torch.cuda.synchronize()
    t2 = time.time()
    print(json.dumps({"case": name, "n": int(s.n_ex + s.n_ey + s.n_hx + s.n_hy), "setup_s": round(t1 - t0, 3),
                      "solve_s": round(t2 - t1, 3), "gammas": [str(g) for g in s.gammas],
                      "residuals": [float(r) for r in s.refined_residuals], "restarts": int(s.refined_restarts),
                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)
