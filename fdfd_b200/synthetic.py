"""Synthetic scattering inputs built directly on the device (bench / multi-GPU slabs).

Same problem family as the reference example scaled per SURVEY.md §8d
(Scattering/example_scattering_by_cylinder.py:8-32: f0 = 10 GHz, 50 cells per wavelength,
eps_r cylinder of radius lambda/2 at the centre, UPML 40 cells with sigma_max = 10, n = 3;
UPML scaling as Scattering_Solver_2D.py:124-153), but evaluated with torch on the GPU for the
rows of one y-slab only, so that a 4096 x 32768 global grid never exists on the host.
These arrays feed K2 directly; they are inputs, not part of any parity claim (parity tests
use the numpy host path of `FDFD2DScatteringSolver`).
"""
from __future__ import annotations

import math

C0 = 299_792_458.0
EPS0 = 8.854187817e-12


def slab_coefficients(Nx, Ny, j0, j1, pol="TE", eps_r=-1e6, f0=10e9, pml=40, sigma_max=10.0, order=3,
                      device="cuda", dtype=None):
    """(ca, cb, cd, sx, sy) for rows [j0, j1) of the Nx x Ny problem; ca = 1/m_a, cb = 1/m_b."""
    import torch

    dtype = dtype or torch.complex128
    lam = C0 / f0
    dx = dy = lam / 50.0
    omega = 2 * math.pi * f0
    k0 = omega / C0
    x = (torch.arange(Nx, device=device, dtype=torch.float64) + 0.5) * dx - Nx * dx / 2
    y = (torch.arange(j0, j1, device=device, dtype=torch.float64) + 0.5) * dy - Ny * dy / 2

    def sigma(idx, n):
        d = torch.minimum(idx, (n - 1) - idx).to(torch.float64)
        s = sigma_max * torch.clamp((pml - d) / pml, min=0.0) ** order
        return s

    sig_x = sigma(torch.arange(Nx, device=device), Nx)
    sig_y = sigma(torch.arange(j0, j1, device=device), Ny)
    Sx = torch.complex(torch.ones_like(sig_x), sig_x / (EPS0 * omega))[None, :]
    Sy = torch.complex(torch.ones_like(sig_y), sig_y / (EPS0 * omega))[:, None]
    inside = (x[None, :] ** 2 + y[:, None] ** 2) <= (0.5 * lam) ** 2
    one = torch.ones((j1 - j0, Nx), device=device, dtype=torch.complex128)
    er = torch.where(inside, torch.full_like(one, complex(eps_r)), one)
    mr = one
    if pol.upper() == "TE":
        ma, mb, dg = mr * (Sx / Sy), mr * (Sy / Sx), er * (Sx * Sy)      # MRyy, MRxx, ERzz
    else:
        ma, mb, dg = er * (Sx / Sy), er * (Sy / Sx), mr * (Sx * Sy)      # ERyy, ERxx, MRzz
    s = 1.0 / (k0 * dx) ** 2
    return (1.0 / ma).to(dtype).contiguous(), (1.0 / mb).to(dtype).contiguous(), dg.to(dtype).contiguous(), s, s


def slab_problem(Nx, Ny, j0, j1, pol="TE", angle_deg=45.0, mask=80, device="cuda", **kw):
    """Slab coefficients plus the plane-wave source and TF/SF mask rows of the same synthetic problem
    (Scattering_Solver_2D.py:89-121, :156-173), all built on the device for rows [j0, j1)."""
    import torch

    ca, cb, cd, sx, sy = slab_coefficients(Nx, Ny, j0, j1, pol=pol, device=device, **kw)
    f0 = kw.get("f0", 10e9)
    lam = C0 / f0
    d = lam / 50.0
    k0 = 2 * math.pi * f0 / C0
    x = (torch.arange(Nx, device=device, dtype=torch.float64) + 0.5) * d - Nx * d / 2
    y = (torch.arange(j0, j1, device=device, dtype=torch.float64) + 0.5) * d - Ny * d / 2
    kx, ky = k0 * math.cos(math.radians(angle_deg)), k0 * math.sin(math.radians(angle_deg))
    ph = -(kx * x[None, :] + ky * y[:, None])
    src = torch.complex(torch.cos(ph), torch.sin(ph))
    jj = torch.arange(j0, j1, device=device)[:, None]
    ii = torch.arange(Nx, device=device)[None, :]
    inside = (ii >= mask) & (ii < Nx - mask) & (jj >= mask) & (jj < Ny - mask)
    q = torch.where(inside, 0.0, 1.0).to(torch.complex128) if 2 * mask < min(Nx, Ny) else torch.ones_like(src)
    return ca, cb, cd, sx, sy, src, q


def sharded_scattering_solve(Nx, Ny, pol="TE", comm=None, rank=0, world=1, device="cuda", pml=40, mask=80,
                             solver=None, repeat=False, **kw):
    """Solve the synthetic scattering problem on `world` y-slabs (one per rank).  Returns
    (x_local (rows, Nx) device tensor, info dict)."""
    import time
    import torch
    from .operators import HelmholtzOperator2D, slab_rows

    def lap(t0):
        torch.cuda.synchronize(device)
        return time.perf_counter() - t0

    j0, j1 = slab_rows(Ny, world, rank)
    t0 = time.perf_counter()
    ca, cb, cd, sx, sy, src, q = slab_problem(Nx, Ny, j0, j1, pol=pol, mask=mask, device=device, pml=pml, **kw)
    t_inputs = lap(t0)
    t0 = time.perf_counter()
    A = HelmholtzOperator2D(pol, Nx, Ny, ca, cb, cd, sx, sy, device=device,
                            rows=(j0, j1) if world > 1 else None, comm=comm)
    s = src.reshape(-1)
    qf = q.reshape(-1)
    b = qf * A.apply(s) - A.apply(qf * s)           # (Q A - A Q) src with two sharded stencil applies
    del src, q, s, qf
    t_operator = lap(t0)
    lam0 = C0 / kw.get("f0", 10e9)
    from .scattering import FDFD2DScatteringSolver
    amp = FDFD2DScatteringSolver.AMPLIFICATION_PER_SQRT_WAVELENGTH * math.sqrt(max(max(Nx, Ny) * (lam0 / 50.0) / lam0, 1.0))
    opts = dict(method="fgmres", precond="mg", rtol=min(1e-8, 1e-6 / (1.25 * amp)), maxit=20000,
                mg_single=1, basis_single=1, line_x_lo=pml, line_x_hi=pml, line_y_lo=pml, line_y_hi=pml)
    opts.update(FDFD2DScatteringSolver.polarisation_defaults(pol))
    opts.update(solver or {})
    t0 = time.perf_counter()
    x, info = A.solve(b, raise_on_noconv=False, **opts)
    info["solve_wall_s"] = lap(t0)               # preconditioner set-up + Krylov solve
    info["precond_setup_s"] = info["solve_wall_s"] - info["seconds"]
    info["inputs_s"], info["operator_rhs_s"], info["rtol"] = t_inputs, t_operator, opts["rtol"]
    if repeat:
        # second right-hand side with the cached preconditioner / Krylov arena (angle sweeps)
        t0 = time.perf_counter()
        x2, info2 = A.solve(b, raise_on_noconv=False, **opts)
        info["repeat_wall_s"], info["repeat_device_s"] = lap(t0), info2["seconds"]
        del x2
    A.close()
    return x.reshape(j1 - j0, Nx), info
