"""Attribute-bag types that form the boundary of the hot path.

These mirror the *names* of ``pyxrd/calculations/data_objects.py:22-271`` in the
reference (``DataObject``, ``AtomTypeData`` ... ``MixtureData``) so that code
written against the reference keeps working.  The calculations in this package
are duck-typed: they accept the reference's own instances just as well as
these (the reference is not importable on the GPU box, hence this mirror).

Extra attributes the reference's model layer attaches ad hoc
(``PhaseData.type/valid_probs/CSDS/G/W/P``, ``SpecimenData.range_theta/
selected_range/z_list``, ``MixtureData.*_mask``) are
listed here as class-level defaults so that a bare instance is well formed.
"""


class DataObject(object):
    """Base bag: every keyword becomes an attribute (reference :22-31)."""

    def __init__(self, **kwargs):
        for name, value in kwargs.items():
            setattr(self, name, value)

    def __repr__(self):  # pragma: no cover - debugging aid
        keys = ", ".join(sorted(self.__dict__))
        return "<%s %s>" % (type(self).__name__, keys)


class AtomTypeData(DataObject):
    par_a = None      # f64[5] Cromer-Mann a_i
    par_b = None      # f64[5] Cromer-Mann b_i
    par_c = None      # constant term
    debye = None      # Debye-Waller B
    charge = 0.0
    weight = 0.0


class AtomData(DataObject):
    atom_type = None  # AtomTypeData
    pn = None         # projected number of atoms at this z
    default_z = None
    z = None          # written by get_factors (z-stretch of interlayer atoms)


class ComponentData(DataObject):
    layer_atoms = None
    interlayer_atoms = None
    volume = None
    weight = None
    d001 = None
    default_c = None
    delta_c = None
    lattice_d = None


class CSDSData(DataObject):
    average = None
    maximum = None
    minimum = None
    alpha_scale = None
    alpha_offset = None
    beta_scale = None
    beta_offset = None


class GonioData(DataObject):
    min_2theta = None
    max_2theta = None
    steps = None
    has_soller1 = False
    soller1 = None
    has_soller2 = False
    soller2 = None
    divergence_mode = "FIXED"
    divergence = None
    mcr_2theta = 0
    has_absorption_correction = None
    absorption = 45.0
    sample_surf_density = 20.0
    radius = None
    wavelength = None
    wavelength_distribution = None
    sample_length = None


class ProbabilityData(DataObject):
    valid = None
    G = None
    W = None
    P = None


class PhaseData(DataObject):
    type = "Phase"
    valid_probs = True
    apply_lpf = True
    apply_correction = True
    components = None
    probability = None
    sigma_star = None
    csds = None
    CSDS = None
    G = None
    W = None
    P = None


class SpecimenData(DataObject):
    goniometer = None
    absorption = None
    phases = None
    range_theta = None
    selected_range = None
    z_list = None
    observed_intensity = None
    total_intensity = None
    scaled_phase_intensities = None
    background_intensity = None
    phase_intensities = None
    correction = None


class MixtureData(DataObject):
    specimens = None
    fractions = None
    bgshifts = None
    scales = None
    fractions_mask = None
    scales_mask = None
    bgshifts_mask = None
    parsed = False
    calculated = False
    optimized = False
    n = 0
    m = 0
