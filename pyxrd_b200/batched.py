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


def get_optimized_residuals(mixtures: Sequence, ctx: Optional[Context] = None):
    """``[get_optimized_residual(m) for m in mixtures]`` (mixture.py:268-276) for a population:
    every mixture is optimised (fractions / scales / background) and left in the same state
    the per-trial callback would leave it in; returns the list of 1-element residual arrays."""
    from .calculations import mixture as cm
    return [cm.get_optimized_residual(m) for m in mixtures]
