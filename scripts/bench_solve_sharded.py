"""Time-to-solution of the synthetic scattering problem on N y-slabs (launch with torchrun for N > 1)."""
import argparse
import datetime
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fdfd_b200  # noqa: E402
from fdfd_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=4096)
    ap.add_argument("--pol", default="TE")
    ap.add_argument("--opts", default="{}")
    ap.add_argument("--repeat", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=600))
        comm = fdfd_b200.SlabComm()
    t0 = time.perf_counter()
    x, info = synthetic.sharded_scattering_solve(args.N, args.N, pol=args.pol, comm=comm, rank=rank, world=world,
                                                 device=dev, solver=eval(args.opts), repeat=args.repeat)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    nrm = torch.linalg.vector_norm(x) ** 2
    if world > 1:
        dist.all_reduce(nrm)
    if rank == 0:
        print(json.dumps(dict(N=args.N, pol=args.pol, n_gpus=world, p2p=bool(comm and comm.peer_mapped),
                              wall_s=round(wall, 3), norm=float(nrm.sqrt()),
                              **{k: (round(v, 4) if isinstance(v, float) and k not in ("relres", "rtol") else v)
                                 for k, v in info.items()})), flush=True)
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
