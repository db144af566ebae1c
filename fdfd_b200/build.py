"""In-tree build of libfdfd_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# Experiment builds live next to the product library and are loaded with FDFD_B200_LIB=<path>:
#   FDFD_COARSE_TRACE=1                 -> libfdfd_b200_trace.so   (phase timings of the persistent coarse solver)
#   FDFD_BUILD_VARIANT="name:-DFLAG=1"  -> libfdfd_b200_<name>.so  (kernel tuning experiments)
_VARIANT = os.environ.get("FDFD_BUILD_VARIANT", "")
_VNAME = "trace" if os.environ.get("FDFD_COARSE_TRACE") else (_VARIANT.split(":", 1)[0] if _VARIANT else "")
LIB = os.path.join(HERE, f"libfdfd_b200_{_VNAME}.so" if _VNAME else "libfdfd_b200.so")
SOURCES = ["api.cu", "yee_assemble.cu", "helm2d.cu", "blas1.cu", "krylov.cu", "comm.cu", "p2p.cu", "mg.cu", "csr_ops.cu", "band.cu", "devpool.cu", "tsqr.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "--fmad=true",
    "-cudart", "static",
]


if _VARIANT and ":" in _VARIANT:
    NVCC_FLAGS.extend(_VARIANT.split(":", 1)[1].split())
if os.environ.get("FDFD_COARSE_TRACE"):
    NVCC_FLAGS.append("-DFDFD_COARSE_TRACE")      # phase timings of the persistent coarse solver (debug)


def _sources():
    out = []
    for s in SOURCES:
        p = os.path.join(CSRC, s)
        if os.path.exists(p):
            out.append(p)
    return out


def _fingerprint(paths):
    h = hashlib.sha256()
    hdrs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "fdfd_b200.h"))
    for p in list(paths) + hdrs:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    stamp = LIB + ".stamp"
    fp = _fingerprint(srcs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, f"build_{_VNAME}" if _VNAME else "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [nvcc, *NVCC_FLAGS, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ldl",
            "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(link, check=True)
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
