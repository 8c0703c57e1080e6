"""Device versions of the hot-path part of ``pyxrd/calculations/goniometer.py``
(LPF and the ADS correction; the theta<->nm axis helpers are GUI utilities and stay
with the reference)."""
from math import sqrt

import numpy as np

from ..data_objects import GonioData, SpecimenData
from ..packing import StaticPack
from ..runtime import default_context


def get_S(soller1, soller2):
    """goniometer.py:15-18 (two scalar operations on constructor arguments)."""
    return sqrt((soller1 * 0.5) ** 2 + (soller2 * 0.5) ** 2), soller1 * soller2


def get_T(range_theta, sigma_star, soller1, soller2):
    """goniometer.py:20-25."""
    theta = np.asarray(range_theta, dtype=float)
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_T(theta, sigma_star, soller1, soller2).reshape(theta.shape)


def get_lorentz_polarisation_factor(range_theta, sigma_star, soller1, soller2, mcr_2theta):
    """goniometer.py:20-34."""
    theta = np.asarray(range_theta, dtype=float)
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_lpf(theta, sigma_star, soller1, soller2, mcr_2theta).reshape(theta.shape)


def get_fixed_to_ads_correction_range(range_theta, goniometer):
    """goniometer.py:36-37: sin(theta), evaluated by the static-table kernel in AUTOMATIC mode."""
    theta = np.asarray(range_theta, dtype=float)
    gon = GonioData(wavelength=1.0, soller1=1.0, soller2=1.0, mcr_2theta=0.0, divergence_mode="AUTOMATIC",
                    divergence=0.0, has_absorption_correction=False, absorption=0.0, sample_surf_density=0.0,
                    sample_length=0.0, radius=1.0, wavelength_distribution=[])
    spec = SpecimenData(goniometer=gon, range_theta=theta.ravel(), selected_range=None, z_list=[0.0],
                        phases=[[]], observed_intensity=np.zeros((theta.size, 1)))
    ctx = default_context()
    with ctx.lock:
        ctx.set_static(StaticPack([spec], {}, n_patterns=1))
        return ctx.read_correction_flat().reshape(theta.shape)
