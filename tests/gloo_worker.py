"""world_size-2 CPU (gloo) worker: host-side sharding logic — slab partition covers the grid without
overlap, and the NCCL unique id generated on rank 0 reaches every rank unchanged through the
torch.distributed broadcast SlabComm uses (communicator creation itself needs GPUs)."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fdfd_b200 import _lib, slab_rows  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1) partition
    for Ny in (7, 64, 4096, 16385):
        j0, j1 = slab_rows(Ny, world, rank)
        t = torch.tensor([j0, j1], dtype=torch.int64)
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, t)
        edges = [int(a[0]) for a in allr] + [int(allr[-1][1])]
        assert edges[0] == 0 and edges[-1] == Ny and all(b > a for a, b in zip(edges, edges[1:])), edges
        assert all(int(allr[r][1]) == int(allr[r + 1][0]) for r in range(world - 1))
        sizes = [int(a[1] - a[0]) for a in allr]
        assert max(sizes) - min(sizes) <= 1
    # 2) unique-id broadcast path
    buf = C.create_string_buffer(128)
    ok = 1
    if rank == 0:
        ok = int(_lib.lib().fdfd_comm_unique_id(buf) == 0)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    okt = torch.tensor([ok])
    dist.broadcast(okt, src=0)
    dist.broadcast(t, src=0)
    if okt.item():
        gathered = [torch.zeros(128, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        assert gathered[0].any()
    # 3) replica distribution of the sweep drivers (fdfd_b200/sweeps.py): slices tile the sweep in rank order and
    #    the gathered rows equal the single-process list; the drivers themselves are exercised with stand-in
    #    solvers (the real ones need a GPU: tests/mgpu_worker.py)
    from fdfd_b200 import sweeps
    import numpy as np
    for n in (1, 2, 5, 20):
        lo, hi = sweeps.replica_slice(n)
        got = sweeps.gather_replicas(list(range(lo, hi)))
        assert got == list(range(n)), (n, got)

    class FakeMode:
        def __init__(self, f):
            self.f = f
        def solve(self):
            self.propagation_constant = np.array([self.f * 1e-9, 2.0])
            self.attenuation_constant = np.array([0.0, 0.5])
    freqs = np.arange(10e9, 17e9, 1e9)
    rows = sweeps.modal_2d_dispersion(freqs, FakeMode, 3)
    assert [r["frequency_GHz"] for r in rows] == [f / 1e9 for f in freqs]
    assert all(abs(r["Mode 1 beta"] - r["frequency_GHz"]) < 1e-9 and np.isnan(r["Mode 3 beta"]) for r in rows)

    class Fake3D:
        def __init__(self, f, sigma):
            self.f, self.sigma = f, sigma
            self.refined_residuals, self.refined_restarts = None, 0
        def solve(self):
            if self.f == 23e9:
                raise RuntimeError("boom")
            self.gammas = np.array([1j * self.f * 1e-10]); self.eigenvalues = self.gammas * 3
    recs = sweeps.periodic_3d_dispersion(np.arange(21e9, 26e9, 1e9), Fake3D, guess_func=lambda f: -1j)
    assert [r["frequency"] for r in recs] == [21e9, 22e9, 23e9, 24e9, 25e9] and "error" in recs[2]
    # a chain starts from guess_func and continues with last_gamma * k0
    assert recs[0]["sigma_guess"] == -1j

    class Fake2DPeriodic:
        def __init__(self, f, guess):
            self.f, self.guess = f, guess
        def solve(self):
            if self.f == 18e9:
                raise RuntimeError("no shifts could be applied")
            self.neff = np.array([0.1 + 1j * self.f * 1e-10, -0.1 - 1j * self.f * 1e-10, 0.3 + 0.2j])
    rows = sweeps.periodic_2d_dispersion(np.arange(16e9, 21e9, 1e9), Fake2DPeriodic, 2, guess_func=lambda f: 0)
    assert [r["Frequency (Hz)"] for r in rows] == [16e9, 17e9, 18e9, 19e9, 20e9]
    assert np.isnan(rows[2]["Alpha_Mode_1"]) and "error" in rows[2] and abs(rows[4]["Beta_Mode_1"] - 2.0) < 1e-9
    assert rows[0]["Alpha_Mode_2"] == -0.1 and "Alpha_Mode_3" not in rows[0]
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK nccl_id=%d" % okt.item(), flush=True)


if __name__ == "__main__":
    main()
