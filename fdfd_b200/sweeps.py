"""Sweep drivers: the callers of the eigen-solvers in three of the four reference configs are
loops over independent problems —
  * the Bloch path of the band diagram       (Band_Diagram_Solver/Band_Diagram_Solver.py:302-323),
  * the frequency loop of the 2-D mode solver (Mode_Solver_2D/Modal_2D_Dispersion.py:28-55),
  * the frequency loop of the 3-D solver      (Periodic_Solver_3D/Periodic_3D_Dispersion.py:36-61).
They do not shard: with one process per GPU (torchrun) every rank takes a contiguous slice of the sweep
("replicas only", no data-path collective) and the results are gathered through ``torch.distributed`` (any
backend; gloo on CPU for the host logic tests).  On one GPU the Bloch path is additionally batched in
lock-step inside ``BandDiagramSolver2D`` (csrc/band.cu, one CTA per path point).
"""
from __future__ import annotations

import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def replica_slice(n_items, world=None, rank=None):
    """Contiguous [lo, hi) of a sweep of `n_items` independent problems for this rank (remainders to low ranks)."""
    d = _dist()
    world = (d.get_world_size() if d else 1) if world is None else int(world)
    rank = (d.get_rank() if d else 0) if rank is None else int(rank)
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_replicas(local):
    """Concatenate the per-rank result lists in rank order on every rank (picklable objects)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return list(local)
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, list(local))
    return [x for p in parts for x in p]


def band_structure_sweep(solver, beta_path, *, num_bands, **kw):
    """`solver.compute_band_structure` with the Bloch path split across the ranks of the process group; every
    rank returns the complete BandStructureResult.  Path points are independent eigen-problems, so the
    result equals the single-process one."""
    from .band_diagram import BandStructureResult
    beta_path = np.asarray(beta_path)
    lo, hi = replica_slice(beta_path.shape[1])
    part = solver.compute_band_structure(beta_path[:, lo:hi], num_bands=num_bands, **kw) if hi > lo else None
    pieces = gather_replicas([None if part is None else (part.frequencies, part.eigenvalues)])
    pols = [p for p in pieces if p is not None]
    freq = {pol: np.concatenate([p[0][pol] for p in pols], axis=1) for pol in pols[0][0]}
    eig = {pol: np.concatenate([p[1][pol] for p in pols], axis=1) for pol in pols[0][1]}
    ticks = solver._tick_positions_from_path(beta_path)
    res = BandStructureResult(beta_path=beta_path, tick_positions=ticks,
                              tick_labels=getattr(solver, "_tick_labels", []) or [""] * len(ticks),
                              frequencies=freq, eigenvalues=eig)
    solver._results = res
    return res


def modal_2d_dispersion(frequencies, build, num_modes, solve_kwargs=None):
    """Rows of the reference's dispersion table (Modal_2D_Dispersion.py:28-55): for every frequency
    ``build(freq)`` returns a configured ModeSolver2D; ``{"frequency_GHz", "Mode m beta", "Mode m alpha"}``
    per frequency, NaN-padded to `num_modes`.  Frequencies are distributed over the ranks."""
    frequencies = np.asarray(frequencies, dtype=float)
    lo, hi = replica_slice(len(frequencies))
    rows = []
    for f in frequencies[lo:hi]:
        s = build(float(f))
        s.solve(**(solve_kwargs or {}))
        row = {"frequency_GHz": f / 1e9}
        n_av = min(num_modes, len(s.propagation_constant), len(s.attenuation_constant))
        for m in range(num_modes):
            row[f"Mode {m + 1} beta"] = float(s.propagation_constant[m]) if m < n_av else np.nan
            row[f"Mode {m + 1} alpha"] = float(s.attenuation_constant[m]) if m < n_av else np.nan
        rows.append(row)
    return gather_replicas(rows)


def periodic_3d_dispersion(frequencies, build, guess_func=None, solve_kwargs=None):
    """The continuation sweep of Periodic_3D_Dispersion.py:36-61: ``build(freq, sigma_guess)`` returns a
    configured PeriodicModeSolver3D; the shift of a frequency is the previous frequency's first gamma times
    k0, or ``guess_func(freq)`` at the start of a chain.  Each rank runs the chain of its own contiguous slice
    (one chain on a single process, exactly the reference's loop).  Returns one dict per frequency."""
    frequencies = np.asarray(frequencies, dtype=float)
    lo, hi = replica_slice(len(frequencies))
    out, last_gamma = [], None
    for f in frequencies[lo:hi]:
        k0 = 2 * np.pi * f / 3e8
        sigma = last_gamma * k0 if last_gamma is not None else (guess_func(f) if guess_func else 0)
        rec = {"frequency": float(f), "sigma_guess": complex(sigma)}
        try:
            s = build(float(f), sigma)
            s.solve(**(solve_kwargs or {}))
            last_gamma = s.gammas[0]
            rec.update(gammas=np.array(s.gammas), eigenvalues=np.array(s.eigenvalues),
                       residuals=None if s.refined_residuals is None else np.array(s.refined_residuals),
                       restarts=s.refined_restarts)
        except Exception as exc:  # noqa: BLE001  (the reference prints a warning and carries on, :60-61)
            rec["error"] = str(exc)
        out.append(rec)
    return gather_replicas(out)


def periodic_2d_dispersion(frequencies, build, num_modes, guess_func=None):
    """The frequency loop of Periodic_2D_Dispersion.py:20-54: ``build(freq, guess)`` returns a configured
    PeriodicModeSolver2D (the reference passes ``guess_func(f)``, 0 by default); per frequency the table row
    ``{"Frequency (Hz)", "Alpha_Mode_m", "Beta_Mode_m"}`` with alpha = Re(neff), beta = Im(neff), NaN where the solve
    failed (the reference prints a warning and carries on).  Frequencies are distributed over the ranks."""
    frequencies = np.asarray(frequencies, dtype=float)
    lo, hi = replica_slice(len(frequencies))
    rows = []
    for f in frequencies[lo:hi]:
        row = {"Frequency (Hz)": float(f)}
        try:
            s = build(float(f), guess_func(f) if guess_func else 0)
            s.solve()
            neff = [complex(s.neff[m]) for m in range(num_modes)]
        except Exception as exc:  # noqa: BLE001
            row["error"] = str(exc)
            neff = [complex(np.nan, np.nan)] * num_modes
        for m in range(num_modes):
            row[f"Alpha_Mode_{m + 1}"] = neff[m].real
            row[f"Beta_Mode_{m + 1}"] = neff[m].imag
        rows.append(row)
    return gather_replicas(rows)
