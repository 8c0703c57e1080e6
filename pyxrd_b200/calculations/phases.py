"""Device versions of ``pyxrd/calculations/phases.py``."""
import logging

import numpy as np

from ..data_objects import GonioData, SpecimenData
from ..packing import BatchPack, StaticPack, collect_atom_types
from ..runtime import default_context

logger = logging.getLogger(__name__)


class _Mix(object):
    pass


def _single_phase_pattern(range_theta, range_stl, soller1, soller2, mcr_2theta, phase, with_lpf):
    theta = np.ascontiguousarray(range_theta, dtype=float).ravel()
    stl = np.ascontiguousarray(range_stl, dtype=float).ravel()
    if theta.size != stl.size:
        raise ValueError("range_theta and range_stl differ in length")
    # divergence_mode "NONE": no machine correction is part of get_intensity (phases.py:81-95)
    gon = GonioData(wavelength=1.0, soller1=soller1, soller2=soller2, mcr_2theta=mcr_2theta, divergence_mode="NONE",
                    divergence=0.0, has_absorption_correction=False, absorption=0.0, sample_surf_density=0.0,
                    sample_length=0.0, radius=1.0, wavelength_distribution=[])
    spec = SpecimenData(goniometer=gon, range_theta=theta, selected_range=None, z_list=[0.0], phases=[[phase]],
                        observed_intensity=np.zeros((theta.size, 1)))
    mix = _Mix()
    mix.specimens, mix.fractions, mix.scales, mix.bgshifts = [spec], np.ones(1), np.ones(1), np.zeros(1)
    saved = phase.apply_lpf
    ctx = default_context()
    with ctx.lock:
        try:
            phase.apply_lpf = bool(with_lpf and saved)
            static = StaticPack([spec], collect_atom_types([spec]), n_patterns=1, stl_override=stl)
            batch = BatchPack([mix], static)
        finally:
            phase.apply_lpf = saved
        ctx.set_static(static)
        ctx.set_batch(batch)
        ctx.compute_intensities()
        return ctx.read_intensities_flat().reshape(np.shape(range_stl))


def get_diffracted_intensity(range_theta, range_stl, phase):
    """phases.py:68-79,106-158: one phase, no LPF.  Invalid probabilities -> zeros."""
    if phase.type == "AbtractPhase":
        raise NotImplementedError
    if phase.type == "Phase" and not phase.valid_probs:
        logger.debug("- calc: get_diffracted_intensity reports: 'Invalid probability found!'")
        return np.zeros_like(np.asarray(range_stl, dtype=float))
    return _single_phase_pattern(range_theta, range_stl, 1.0, 1.0, 0.0, phase, with_lpf=False)


def get_intensity(range_theta, range_stl, soller1, soller2, mcr_2theta, phase):
    """phases.py:81-95: one phase with the Lorentz-polarisation factor (if ``apply_lpf``)."""
    if phase.type == "AbtractPhase":
        raise NotImplementedError
    if phase.type == "Phase" and not phase.valid_probs:
        return np.zeros_like(np.asarray(range_stl, dtype=float))
    return _single_phase_pattern(range_theta, range_stl, soller1, soller2, mcr_2theta, phase, with_lpf=True)


def get_structure_factors(range_stl, G, comp_list):
    """phases.py:20-39: structure factors and phase factors of the components, c128[X, G] each
    (``components.get_factors`` per component, evaluated on the device)."""
    from .components import get_factors
    range_stl = np.asarray(range_stl, dtype=float)
    shape = range_stl.shape + (G,)
    SF = np.zeros(shape, dtype=np.complex128)
    PF = np.zeros(shape, dtype=np.complex128)
    for i, comp in enumerate(comp_list):
        SF[:, i], PF[:, i] = get_factors(range_stl, comp)
    return SF, PF


def get_Q_matrices(Q, CSDS_max):
    """phases.py:41-46: ``Qn[n] = Q ** (n + 1)`` for n = 0 .. CSDS_max by repeated ``mmult`` (math_tools.py:16-17),
    complex128[CSDS_max + 1, X, r, r].  The batch path never forms this stack; this is the
    reference's helper itself, on the device (``k_leaf_matrix_powers``, numpy's rounding of the products and sums)."""
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_matrix_powers(Q, CSDS_max)


def get_absolute_scale(components, CSDS_real_mean, W):
    """phases.py:48-66: a dozen scalar operations on model constants (inside the batch path the prologue kernel does
    them per phase, k_phase_prologue); first G diagonal entries of W, ``None`` components skipped, 0 if the mass is 0."""
    W = np.diag(np.asarray(W, dtype=float))
    mean_volume = mean_d001 = mean_density = 0.0
    for i, comp in enumerate(components):
        if comp is not None:
            mean_volume += comp.volume * W[i]
            mean_d001 += comp.d001 * W[i]
            mean_density += (comp.weight * W[i] / comp.volume)
        else:
            logger.debug("- calc: get_absolute_scale reports: 'Zero observations found!'")
    mean_mass = (CSDS_real_mean * mean_volume ** 2 * mean_density)
    return mean_d001 / mean_mass if mean_mass != 0.0 else 0.0
