"""Whole-population entry points (additive API, SURVEY.md §8(b) "New API"): evaluate K
candidate mixtures in one launch sequence instead of K calls of the per-trial callback."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import settings
from .packing import BatchPack, StaticPack, collect_atom_types
from .runtime import Context, default_context


def pack_population(mixtures: Sequence, static: Optional[StaticPack] = None):
    """(StaticPack, BatchPack) of K candidates that share specimens (theta grids, goniometer,
    observed patterns); the static tables are taken from the first candidate."""
    if not len(mixtures):
        raise ValueError("empty population")
    if static is None:
        types = {}
        for mix in mixtures:
            for key in collect_atom_types(mix.specimens):
                types.setdefault(key, len(types))
        first = mixtures[0]
        z = 0
        for spec in first.specimens:
            if spec is not None:
                z = len(spec.z_list)
                break
        static = StaticPack(first.specimens, types, n_patterns=z)
    batch = BatchPack(mixtures, static, bgshift_enabled=bool(settings.get("BGSHIFT")))
    return static, batch


def population_residuals(mixtures: Sequence, ctx: Optional[Context] = None, want_intensities: bool = False):
    """residuals f64[K, 1+S] of K candidates at their current fractions / scales / bgshifts
    (``calculate_mixture`` semantics, mixture.py:222-257), one C-ABI call."""
    ctx = ctx or default_context()
    static, batch = pack_population(mixtures)
    with ctx.lock:
        ctx.set_static(static)
        ctx.set_residual_method(settings.get("RESIDUAL_METHOD"))
        res, inten = ctx.evaluate_host(batch, want_intensities=want_intensities)
    return (res, inten) if want_intensities else res


def population_phase_intensities(mixtures: Sequence, ctx: Optional[Context] = None):
    """list over candidates of list over specimens of f64[M, Z, X_s] (``calculate_phase_intensities``)."""
    ctx = ctx or default_context()
    static, batch = pack_population(mixtures)
    with ctx.lock:
        ctx.set_static(static)
        ctx.set_batch(batch)
        ctx.compute_intensities()
        return ctx.split_per_specimen(ctx.read_intensities_flat(), (batch.M, static.Z))


def _device_optimisable(mix) -> bool:
    """the device optimiser covers the reference's default masks: every fraction and every scale
    free, background shifts all free or all fixed (mixture/models/mixture.py:64-73)"""
    m_all = len(np.asarray(mix.fractions).ravel())
    n_all = len(mix.specimens)
    return (settings.get("RESIDUAL_METHOD") == "Rp"        # the device optimiser minimises Rp only
            and mix.m == m_all and mix.n == n_all and mix.o in (0, n_all) and m_all <= 11 and n_all <= 8
            and all(sp is not None and np.size(sp.observed_intensity) > 0 for sp in mix.specimens))


def get_optimized_residuals(mixtures: Sequence, ctx: Optional[Context] = None, finalize: bool = False,
                            outer: int = 0, newton: int = 0):
    """``[get_optimized_residual(m) for m in mixtures]`` (mixture.py:268-276) for a whole
    population in one launch sequence: phase intensities, then the device-resident optimiser
    (majorise-minimise on the Rp objective, ``pyxrd_optimize``) for every candidate at once.

    Every mixture is left as ``optimize_mixture`` leaves it (fractions / scales rounded to 6
    digits, ``residual``, ``optimized``); with ``finalize=True`` the concluding
    ``calculate_mixture`` of the reference is run too (patterns + ``residuals`` on the host).
    Returns the list of 1-element residual arrays.  Contract for the optimum: <= the reference's
    L-BFGS-B end point * (1 + 5e-3); mixtures with partially fixed variables go
    through the scipy-driven ``optimize_mixture`` with the device evaluating the objective."""
    from .calculations import mixture as cm
    ctx = ctx or default_context()
    out = [None] * len(mixtures)
    todo = []
    for i, mix in enumerate(mixtures):
        if mix.optimized:
            out[i] = cm.get_optimized_residual(mix)
            continue
        try:
            if not mix.parsed:
                cm._parse_bookkeeping(mix)
        except AssertionError:
            if settings.get("DEBUG"):
                raise
            out[i] = cm.get_optimized_residual(mix)        # returns the mixture's state unchanged
            continue
        if _device_optimisable(mix):
            todo.append(i)
        else:
            mix.parsed = False
            mix.m, mix.n = mix.real_m, mix.real_n
            out[i] = cm.get_optimized_residual(mix)
    # one device batch per value of o: the candidates of a batch share "background refined or not"
    for free_bg in (True, False):
        rows = [i for i in todo if (mixtures[i].o > 0) == free_bg]
        if not rows:
            continue
        group = [mixtures[i] for i in rows]
        static, batch = pack_population(group)
        refine_bg = free_bg and bool(settings.get("BGSHIFT"))
        with ctx.lock:
            ctx.set_static(static)
            ctx.set_batch(batch)
            ctx.compute_intensities()
            ctx.optimize(refine_bg=refine_bg, outer=outer, newton=newton)
            sol = ctx.read_solutions()
            res = ctx.read_opt_residuals()
        M, S = batch.M, batch.S
        for row, i in enumerate(rows):
            mix = mixtures[i]
            mix.fractions = np.around(sol[row, :M] / np.sum(sol[row, :M]), 6)
            mix.scales = np.array(sol[row, M:M + S]).round(6)
            mix.bgshifts = np.array(sol[row, M + S:M + 2 * S]) if free_bg else np.asarray(mix.bgshifts, float)
            mix.residual = np.array([res[row, 0]])
            mix.parsed = False            # phase intensities were not copied to the host
            mix.m, mix.n = mix.real_m, mix.real_n
            mix.calculated = False
            mix.optimized = True
            out[i] = mix.residual
    if todo:
        if finalize:
            for i in todo:
                cm.calculate_mixture(mixtures[i])
    return out
