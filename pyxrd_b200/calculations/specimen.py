"""Device versions of ``pyxrd/calculations/specimen.py``."""
import numpy as np

from ..packing import BatchPack, StaticPack, collect_atom_types
from ..runtime import default_context


class _Mix(object):
    pass


def _solo_mixture(specimen, n_phases, fractions=None, scale=1.0, bgshift=0.0):
    mix = _Mix()
    mix.specimens = [specimen]
    mix.fractions = np.ones(n_phases) if fractions is None else np.asarray(fractions, dtype=float)
    mix.scales = np.array([float(scale)])
    mix.bgshifts = np.array([float(bgshift)])
    return mix


def get_clipped_intensities(specimen):
    """specimen.py:16-19 (pure indexing of host arrays the caller owns)."""
    exp = specimen.observed_intensity.transpose()[..., specimen.selected_range]
    cal = specimen.total_intensity[..., specimen.selected_range]
    return exp, cal


def get_machine_correction_range(specimen):
    """specimen.py:21-43."""
    ctx = default_context()
    with ctx.lock:
        ctx.set_static(StaticPack([specimen], collect_atom_types([specimen])))
        return ctx.read_correction_flat()


def calculate_phase_intensities(specimen):
    """specimen.py:45-82 -> (correction f64[X], phase intensities f64[M, Z, X])."""
    M = len(specimen.phases[0]) if len(specimen.phases) else 0
    Z = len(specimen.z_list)
    ctx = default_context()
    with ctx.lock:
        static = StaticPack([specimen], collect_atom_types([specimen]))
        ctx.set_static(static)
        corr = ctx.read_correction_flat()
        if M == 0:
            return corr, np.zeros((0, Z, corr.size))
        batch = BatchPack([_solo_mixture(specimen, M)], static)
        ctx.set_batch(batch)
        ctx.compute_intensities()
        flat = ctx.read_intensities_flat()
    return corr, flat.reshape(M, Z, corr.size)


def calculate_scaled_intensities(specimen, scale, fractions, bgshift):
    """specimen.py:84-88: needs ``specimen.correction`` / ``.phase_intensities`` (set by
    parse_mixture); sets background / scaled / total intensities on the specimen."""
    pi = np.ascontiguousarray(specimen.phase_intensities, dtype=float)
    M, Z, X = pi.shape
    fractions = np.asarray(fractions, dtype=float).ravel()
    shadow = type(specimen).__new__(type(specimen))
    shadow.__dict__.update(specimen.__dict__)
    shadow.phases = [[None] * M for _ in range(Z)]          # patterns are supplied, not recomputed
    ctx = default_context()
    with ctx.lock:
        static = StaticPack([shadow], {}, n_patterns=Z)
        ctx.set_static(static)
        batch = BatchPack([_solo_mixture(shadow, M, fractions, scale, bgshift)], static)
        ctx.set_batch(batch)
        ctx.set_intensities(pi)
        ctx.set_residual_method("Rp")                    # only the scaled / total patterns are read back
        ctx.compute_residuals(want_scaled=True)
        specimen.scaled_phase_intensities = ctx.read_scaled_flat().reshape(M, Z, X)
        specimen.total_intensity = ctx.read_totals_flat().reshape(Z, X)
    # bgshift * correction with the caller's own correction array, like the reference
    specimen.background_intensity = specimen.total_intensity[0] * 0.0 + bgshift * np.asarray(specimen.correction)
    return specimen
