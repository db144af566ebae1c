"""(Named test_zz_* so that these multi-process tests run after the single-GPU parity tests.)
Multi-rank tests: world_size-2 gloo on CPU for the host-side sharding logic, and (gpu, needs
>= 2 devices) NCCL y-slab sharding of the operator and the Krylov solve against the single-domain
result."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(script, nproc, timeout=240, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", script)]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=e)


def test_gloo_world2_host_logic():
    r = _torchrun("gloo_worker.py", 2)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["p2p", "nccl"])
def test_slab_sharding(path):
    """Sharded apply / Krylov / multigrid / drop-in class against the single-domain result, on the peer-memory
    data path (ghost rows pushed over NVLink by the stencil kernel itself, flag-based all-gather / all-reduce
    kernels) and on the NCCL fallback path (FDFD_P2P=0)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun("mgpu_worker.py", min(n, 4), env={"FDFD_P2P": "1" if path == "p2p" else "0", "FDFD_EXPECT_PATH": path})
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
