"""Multi-rank worker (launched by torchrun from tests/test_zz_multigpu.py): y-slab sharded operator
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
    expect = os.environ.get("FDFD_EXPECT_PATH")
    if expect and comm.peer_mapped != (expect == "p2p"):
        print(f"rank {rank}: data path is {'p2p' if comm.peer_mapped else 'nccl'}, expected {expect}", flush=True)
        sys.exit(1)
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
    # sharded multigrid preconditioner == single-domain multigrid (same algorithm; the line solver only
    # partitions its chunks differently), then a full MG-preconditioned scattering solve
    sys.path.insert(0, ROOT)
    from oracle import scattering_oracle as so
    for pol in ("TE", "TM"):
        N, pw = 256, 12
        p = so.cylinder_problem(N, eps_r=4.0, pml_width=pw, mask=pw + 8)
        ca, cb, cd, sx, sy = so.stencil_coefficients(p, pol)
        v = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
        j0, j1 = fdfd_b200.slab_rows(N, world, rank)
        lines = dict(line_x_lo=pw, line_x_hi=pw, line_y_lo=pw, line_y_hi=pw, mg_coarse_its=6, mg_shift_im=0.5, mg_omega=0.8)
        full = fdfd_b200.HelmholtzOperator2D(pol, N, N, ca, cb, cd, sx, sy, device=dev)
        zf = full.precond_apply(torch.from_numpy(v).to(dev), precond="mg", **lines).reshape(N, N)
        part = fdfd_b200.HelmholtzOperator2D(pol, N, N, ca[j0:j1], cb[j0:j1], cd[j0:j1], sx, sy, device=dev,
                                             rows=(j0, j1), comm=comm)
        # coarsest level replicated on every rank (default for small coarse grids) and distributed (what large
        # grids use: with a peer-mapped heap the coarse solve is ONE persistent kernel per rank exchanging ghost
        # rows, inner products and line interface values over NVLink; on the NCCL path a launch chain)
        for rep in (-1, 0):
            zp = part.precond_apply(torch.from_numpy(np.ascontiguousarray(v[j0:j1])).to(dev), precond="mg",
                                    mg_coarse_replicate=rep, **lines)
            err = (torch.linalg.vector_norm(zp.reshape(-1) - zf[j0:j1].reshape(-1)) / torch.linalg.vector_norm(zf[j0:j1])).item()
            if not err <= 1e-7:
                failures.append(f"mg precond {pol} (coarse_replicate={rep}): sharded vs full rel diff {err}")
        xd, inf_d = part.solve(torch.from_numpy(np.ascontiguousarray(so.rhs(so.build_A(p, pol), p).reshape(N, N)[j0:j1])).to(dev).reshape(-1),
                               method="fgmres", precond="mg", rtol=1e-9, maxit=400, mg_coarse_replicate=0, mg_single=1,
                               line_x_lo=pw, line_x_hi=pw, line_y_lo=pw, line_y_hi=pw)
        if inf_d["relres"] > 1e-9:
            failures.append(f"mg solve {pol} with the distributed coarse level: relres {inf_d['relres']}")
        # solve
        A = so.build_A(p, pol)
        b = so.rhs(A, p).reshape(N, N)
        xf, inf_f = full.solve(torch.from_numpy(b).to(dev).reshape(-1), method="fgmres", precond="mg", rtol=1e-9,
                               maxit=400, line_x_lo=pw, line_x_hi=pw, line_y_lo=pw, line_y_hi=pw)
        xp, inf_p = part.solve(torch.from_numpy(np.ascontiguousarray(b[j0:j1])).to(dev).reshape(-1), method="fgmres",
                               precond="mg", rtol=1e-9, maxit=400, line_x_lo=pw, line_x_hi=pw, line_y_lo=pw, line_y_hi=pw)
        ref = xf.reshape(N, N)[j0:j1].reshape(-1)
        err = (torch.linalg.vector_norm(xp - ref) / torch.linalg.vector_norm(ref)).item()
        if err > 1e-6 or inf_p["relres"] > 1e-9 or abs(inf_p["iterations"] - inf_f["iterations"]) > 0.25 * inf_f["iterations"]:
            failures.append(f"mg solve {pol}: err {err}, its {inf_p['iterations']} vs {inf_f['iterations']}, relres {inf_p['relres']}")
        if rank == 0:
            print(f"sharded MG {pol}: precond ok, solve its {inf_p['iterations']} (single {inf_f['iterations']})", flush=True)
        part.close()
        full.close()
    # drop-in class, sharded: every rank gets the full field, equal to the single-GPU result
    N = 256
    f0 = 10e9
    lam = 299792458 / f0
    fields = []
    for c in (None, comm):
        sim = fdfd_b200.FDFD2DScatteringSolver(f0, N / 50 * lam, N / 50 * lam, N, N, device=dev, comm=c)
        sim.add_object(4.0, 1.0, (sim.X ** 2 + sim.Y ** 2) <= (0.5 * lam) ** 2)
        sim.add_UPML(pml_width=16, sigma_max=10)
        sim.add_mask(30)
        sim.add_source("plane_wave", angle_deg=45.0)
        fields.append(sim.solve_total_field_TM())
        if c is None:
            sim.close()
        # the sharded solver stays open on purpose: comm.close() below must release its cached
        # preconditioner (captured NCCL graph) before destroying the communicator
    err = np.linalg.norm(fields[1] - fields[0]) / np.linalg.norm(fields[0])
    if fields[1].shape != (N, N) or err > 1e-6:
        failures.append(f"sharded FDFD2DScatteringSolver: rel diff {err}")
    # replica distribution of a Bloch sweep (fdfd_b200.sweeps): every rank solves a slice of the path on its own
    # GPU; the gathered band structure equals the single-process one (independent eigen-problems)
    bs = fdfd_b200.BandDiagramSolver2D(a=1.0, Nx=16, background_er=9.0)
    bs.device = dev
    bs.add_circular_inclusion(radius=0.3, er=1.0)
    bp = np.array([[0.2, 1.0, 2.0, 3.0, 0.7], [0.1, 0.0, 1.5, 3.1, 2.2]])
    with torch.cuda.device(dev):
        single = bs.compute_band_structure(bp, num_bands=3)
        spread = fdfd_b200.sweeps.band_structure_sweep(bs, bp, num_bands=3)
    for pol in ("TE", "TM"):
        d = np.max(np.abs(single.eigenvalues[pol] - spread.eigenvalues[pol]))
        if spread.eigenvalues[pol].shape != (3, 5) or d > 1e-9 * np.max(np.abs(single.eigenvalues[pol])):
            failures.append(f"distributed Bloch sweep {pol}: max diff {d}")
    flag = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES:", failures, flush=True)
    comm.close()
    # operators of a closed communicator refuse to run instead of touching freed peer memory
    try:
        op_closed = sim.operator("TM"); op_closed.apply(torch.zeros(op_closed.n, dtype=torch.complex128, device=dev))
        failures.append("apply on a closed sharded operator did not raise")
    except (fdfd_b200.FdfdError, ValueError, KeyError):
        pass
    if failures:
        flag += 1
    dist.destroy_process_group()
    if flag.item():
        sys.exit(1)
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
