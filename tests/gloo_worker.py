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
    dist.destroy_process_group()
    if rank == 0:
        print("GLOO_OK nccl_id=%d" % okt.item(), flush=True)


if __name__ == "__main__":
    main()
