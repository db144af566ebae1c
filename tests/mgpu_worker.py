"""Multi-rank worker (launched by torchrun from tests/test_multigpu.py): y-slab sharded operator
and Krylov solve must reproduce the single-domain result on every rank's slab."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fdfd_b200  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    comm = fdfd_b200.SlabComm()
    rng = np.random.default_rng(42)           # same stream on every rank -> same global arrays
    failures = []
    for (Nx, Ny), pol, bc in [((96, 70), "TE", (0, 0)), ((96, 70), "TM", (0, 0)), ((64, 64), "TE", (1, 1)),
                              ((64, 64), "TM", (0, 1)), ((130, 33), "TM", (1, 0)), ((512, 512), "TE", (0, 0))]:
        ca = rng.uniform(0.5, 2, (Ny, Nx)) + 1j * rng.uniform(-1, 1, (Ny, Nx))
        cb = rng.uniform(0.5, 2, (Ny, Nx)) + 1j * rng.uniform(-1, 1, (Ny, Nx))
        cd = rng.uniform(0.5, 2, (Ny, Nx)) + 1j * rng.uniform(-1, 1, (Ny, Nx))
        x = rng.standard_normal((Ny, Nx)) + 1j * rng.standard_normal((Ny, Nx))
        kinc = rng.standard_normal(2)
        phase = (np.exp(-1j * kinc[0] * Nx), np.exp(-1j * kinc[1] * Ny))
        full = fdfd_b200.HelmholtzOperator2D(pol, Nx, Ny, ca, cb, cd, 3.3, 2.7, bc=bc, phase=phase, device=dev)
        yfull = full.apply(torch.from_numpy(x).to(dev)).reshape(Ny, Nx)
        j0, j1 = fdfd_b200.slab_rows(Ny, world, rank)
        part = fdfd_b200.HelmholtzOperator2D(pol, Nx, Ny, ca[j0:j1], cb[j0:j1], cd[j0:j1], 3.3, 2.7, bc=bc,
                                             phase=phase, device=dev, rows=(j0, j1), comm=comm)
        ypart = part.apply(torch.from_numpy(np.ascontiguousarray(x[j0:j1])).to(dev)).reshape(j1 - j0, Nx)
        if not torch.equal(ypart, yfull[j0:j1]):
            err = (ypart - yfull[j0:j1]).abs().max().item()
            failures.append(f"apply {pol} {bc} {Nx}x{Ny}: max diff {err}")
        part.close()
        full.close()
    # sharded Krylov solve (lossy, well conditioned) == single-domain solve
    N = 96
    ca = 1.0 + 0.1 * rng.standard_normal((N, N)) + 0.05j
    cb = 1.0 + 0.1 * rng.standard_normal((N, N)) - 0.05j
    cd = (-300.0 + 40.0j) * np.ones((N, N))
    b = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    for method in ("bicgstab", "fgmres"):
        full = fdfd_b200.HelmholtzOperator2D("TM", N, N, ca, cb, cd, 63.0, 63.0, device=dev)
        xf, _ = full.solve(torch.from_numpy(b).to(dev).reshape(-1), method=method, precond="jacobi", rtol=1e-11)
        j0, j1 = fdfd_b200.slab_rows(N, world, rank)
        part = fdfd_b200.HelmholtzOperator2D("TM", N, N, ca[j0:j1], cb[j0:j1], cd[j0:j1], 63.0, 63.0, device=dev,
                                             rows=(j0, j1), comm=comm)
        xp, info = part.solve(torch.from_numpy(np.ascontiguousarray(b[j0:j1])).to(dev).reshape(-1), method=method,
                              precond="jacobi", rtol=1e-11)
        ref = xf.reshape(N, N)[j0:j1].reshape(-1)
        err = (torch.linalg.vector_norm(xp - ref) / torch.linalg.vector_norm(ref)).item()
        if err > 1e-9 or info["relres"] > 1e-11:
            failures.append(f"solve {method}: err {err}, relres {info['relres']}")
        part.close()
        full.close()
    flag = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES:", failures, flush=True)
    comm.close()
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
